"""Summarise an .ncu-rep (raw page) into a markdown table of the metrics the roofline needs."""
import csv, subprocess, sys
KEYS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("lts__t_bytes.sum", "L2 bytes"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1 %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
        ("sm__inst_executed_pipe_tensor.sum", "tensor inst"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__shared_mem_per_block_dynamic", "dyn smem"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_sb"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle"),
        ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %")]
def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print("| kernel | " + " | ".join(n for _, n in KEYS) + " |")
    print("|---|" + "---|" * len(KEYS))
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        name = d.get("Kernel Name", "?").replace("void ", "").replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
        name = name.split("(")[0][:48]
        cells = []
        for k, _ in KEYS:
            if k in d:
                u = units[hdr.index(k)]
                cells.append(("%s %s" % (d[k], u)).strip())
            else:
                cells.append("-")
        print("| " + name + " | " + " | ".join(cells) + " |")
if __name__ == "__main__":
    main(sys.argv[1])
