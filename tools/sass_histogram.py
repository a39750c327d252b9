"""SASS evidence of what the built library contains (cuobjdump -sass of libddf_b200.so): per kernel the counts of the
instructions that prove tcgen05 / TMEM / TMA / bulk-copy / vector reductions.  Written to profiles/ by hand:
    python tools/sass_histogram.py > profiles/r2_sass_histogram.md"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "3d-dual-fusion_b200", "libddf_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ["ACQBULK", "PREEXIT", "UTCHMMA", "UTCBAR", "LDTM", "UTCATOM", "UTMALDG", "UBLKCP", "LDGSTS", "SYNCS", "REDG", "RED.", "ATOMG", "ATOMS", "HMMA", "FFMA", "DFMA", "DADD", "SHFL", "LDG", "STG", "LDS", "STS"]
cur = None
hist = collections.OrderedDict()
variants = collections.Counter()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::", "", name)
        name = name.split("(")[0].replace("void ", "")
        cur = hist.setdefault(name, collections.Counter())
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    cur["_total"] += 1
    for k in KEYS:
        if op.startswith(k):
            cur[k] += 1
            if k in ("UTMALDG", "UTCHMMA", "UBLKCP", "REDG", "RED."):
                variants[op] += 1
            break
print("# SASS instruction histogram of libddf_b200.so (sm_100a), per kernel\n")
print("`cuobjdump -sass`; UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM = tcgen05.ld, UTMALDG = TMA tensor load "
      "(.GATHER4 = gather4), UBLKCP = cp.async.bulk, LDGSTS = cp.async, SYNCS = mbarrier, REDG / RED = red.global.\n")
cols = ["UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UBLKCP", "LDGSTS", "SYNCS", "REDG", "RED.", "ATOMG", "SHFL", "FFMA", "DFMA"]
print("| kernel | instr | " + " | ".join(cols) + " |")
print("|---|---:|" + "---:|" * len(cols))
tot = collections.Counter()
for name, c in hist.items():
    if not any(c[k] for k in ("UTCHMMA", "UTMALDG", "UBLKCP", "LDGSTS", "REDG", "RED.", "LDTM")) and c["_total"] < 400:
        continue
    print("| %s | %d | " % (name[:70], c["_total"]) + " | ".join(str(c[k]) if c[k] else "" for k in cols) + " |")
    tot.update(c)
print("\nTotals over the listed kernels: " + ", ".join("%s %d" % (k, tot[k]) for k in cols if tot[k]))
print("\nVariants: " + ", ".join("%s x%d" % kv for kv in sorted(variants.items())))
print("\n%d kernels in the library." % len(hist))
n_wait = sum(1 for c in hist.values() if c["ACQBULK"])
n_trig = sum(1 for c in hist.values() if c["PREEXIT"])
print("\nProgrammatic dependent launch: ACQBULK (griddepcontrol.wait) in %d of %d kernels, PREEXIT "
      "(griddepcontrol.launch_dependents) in %d." % (n_wait, len(hist), n_trig))
