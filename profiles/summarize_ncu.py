"""Summarise an `ncu --page raw --csv` export into a compact table (+ JSON) of the metrics the roofline uses."""
import csv
import json
import sys

METRICS = [
    ("us", "gpu__time_duration.sum"), ("grid", "launch__grid_size"), ("block", "launch__block_size"),
    ("regs", "launch__registers_per_thread"), ("dyn_smem_B", "launch__shared_mem_per_block_dynamic"),
    ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("sm_thru_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("dram_thru_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("dram_read_MB", "dram__bytes_read.sum"), ("dram_write_MB", "dram__bytes_write.sum"),
    ("warp_inst", "smsp__inst_executed.sum"),
    ("fma_pct", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
    ("alu_pct", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
    ("fp64_pct", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
    ("lsu_pct", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
    ("tensor_pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    ("stall_barrier", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
    ("stall_long_sb", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
    ("stall_short_sb", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
    ("stall_wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
    ("stall_mio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"),
    ("stall_math", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"),
    ("smem_bank_conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main(path, out_json=None, labels_json=None):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    labels = json.load(open(labels_json)) if labels_json else []
    summary = {}
    seen = {}
    for li, r in enumerate(rows[2:]):
        name = r[idx["Kernel Name"]].split("(")[0].replace("ovdet::", "").replace("void ", "")
        seen[name] = seen.get(name, 0) + 1
        key = name if seen[name] == 1 else "%s#%d" % (name, seen[name])
        if li < len(labels):        # profiles/prof_driver.py's label of this launch, then the kernel name
            key = "%s | %s" % (labels[li], name)
        d = {}
        for short, m in METRICS:
            if m not in idx:
                continue
            v, u = r[idx[m]], units[idx[m]]
            if short.endswith("_MB"):
                d[short] = round(to_bytes(v, u) / 1e6, 3)
            elif short == "us":
                f = float(v.replace(",", ""))
                d[short] = round(f / 1e3 if u in ("nsecond", "ns") else (f * 1e3 if u in ("msecond", "ms") else f), 2)
            else:
                try:
                    d[short] = round(float(v.replace(",", "")), 3)
                except ValueError:
                    d[short] = v
        d["dram_bytes_per_launch"] = int((d.get("dram_read_MB", 0) + d.get("dram_write_MB", 0)) * 1e6)
        summary[key] = d
    cols = ["us", "grid", "regs", "warps_active_pct", "issue_active_pct", "dram_thru_pct", "dram_read_MB", "dram_write_MB",
            "warp_inst", "fma_pct", "alu_pct", "fp64_pct", "lsu_pct", "tensor_pct", "stall_barrier", "stall_long_sb",
            "stall_short_sb", "stall_wait", "stall_mio"]
    print("| kernel | " + " | ".join(cols) + " |")
    print("|---|" + "---|" * len(cols))
    for k, d in summary.items():
        print("| %s | " % k + " | ".join(str(d.get(c, "")) for c in cols) + " |")
    if out_json:
        json.dump(summary, open(out_json, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None, sys.argv[3] if len(sys.argv) > 3 else None)
