"""Turn the files a tools/gpu_profile.sh run leaves in gpurun_out/ into the markdown summaries kept under profiles/.

    python tools/profile_report.py launches  gpurun_out/launches.csv  gpurun_out/bench.json  > profiles/<name>.md
    python tools/profile_report.py kernels   gpurun_out/prof_spconv.raw.csv [more.raw.csv ...] > profiles/<name>.md
"""
import collections
import csv
import json
import re
import sys

OURS = ("spconv", "msda", "subm_", "conv_", "pairs_", "ddf::", "vox", "round_tf32", "dense_", "transpose_filters",
        "fill_i32", "fps", "ball", "group", "gather", "bn_", "hash_", "popc", "scan", "xty_", "bigate_", "bias_relu_",
        "relu_dropout_", "add_dropout_", "col_sum", "split_", "local_attn", "project_assign", "first_occ", "index_rows",
        "scatter_first", "bev_", "sparse_to", "mark_", "compact_")

KEYS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1TEX %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_sb"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_sb")]


def short(name):
    name = re.sub(r"\(.*", "", name).replace("void ", "").replace("(anonymous namespace)::", "")
    return name.replace("<unnamed>::", "").replace("at::native::", "native::")[:80]


def launches(path, bench_path=None):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot, cnt = {}, collections.Counter()
    body = rows[hi + 1:]
    if len(sys.argv) > 4:      # "laststep": the launches between the last two optimizer launches = one warm step
        marks = [i for i, r in enumerate(body) if len(r) > kn and "adam" in r[kn].lower()]
        ends = [m for i, m in enumerate(marks) if i + 1 == len(marks) or marks[i + 1] - m > 50]   # last launch of each step
        if len(ends) >= 2:
            body = body[ends[-2] + 1:ends[-1] + 1]
    for r in body:
        if len(r) <= mv:
            continue
        n = short(r[kn])
        tot[n] = tot.get(n, 0.0) + float(r[mv].replace(",", "")) / 1000.0
        cnt[n] += 1
    T = sum(tot.values())
    print("Cold-cache, serialised per-launch times: compare SHARES, not absolutes.\n")
    if bench_path:
        b = json.load(open(bench_path))
        print("Live bench line of the same build (not under ncu): %.2f samples/s device-resident (%.1f ms/step), e2e %.2f "
              "samples/s; conv fwd+dgrad %.0f%% of the step at %.1f TFLOP/s; wgrad %.1f ms/step at %.1f TFLOP/s; "
              "deform-attn fwd %.0f GB/s, bwd %.0f GB/s (algorithmic bytes / event time).\n" % (
                  b["value"], b["ms_per_step"], b["e2e"]["value"], 100 * b["roofline"]["share_of_step"],
                  b["roofline"]["achieved"], b["kernels"]["sparse_conv_wgrad"]["ms_per_step"],
                  b["kernels"]["sparse_conv_wgrad"]["achieved_TFLOPs"], b["kernels"]["deform_attn_fwd"]["achieved_GBs"],
                  b["kernels"]["deform_attn_bwd"]["achieved_GBs"]))
    ours = sum(v for k, v in tot.items() if k.startswith(OURS))
    print("Kernels of this library: %.1f%% of the listed GPU time; the rest is PyTorch (cuBLAS GEMMs, LayerNorm, dropout, "
          "elementwise and reductions of the fusion encoder, optimizer).\n" % (100 * ours / T))
    print("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
    for k, v in sorted(tot.items(), key=lambda x: -x[1])[:45]:
        print("| %s | %d | %.1f | %.1f%% |" % (k, cnt[k], v, 100 * v / T))
    print("\nTotal %.1f us over %d launches." % (T, sum(cnt.values())))


def kernels(paths):
    for f in paths:
        rows = list(csv.reader(open(f)))
        hdr, units = rows[0], rows[1]
        print("## %s\n" % f.split("/")[-1])
        print("| kernel | " + " | ".join(n for _, n in KEYS) + " |")
        print("|---|" + "---|" * len(KEYS))
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            cells = []
            for k, _ in KEYS:
                v = d.get(k, "-")
                u = units[hdr.index(k)] if k in hdr else ""
                try:
                    v = "%.3g" % float(v)
                except ValueError:
                    pass
                cells.append((v + " " + u).strip())
            print("| " + short(d["Kernel Name"]) + " | " + " | ".join(cells) + " |")
        print()


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
    else:
        kernels(sys.argv[2:])
