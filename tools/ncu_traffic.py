"""Turns an `ncu --set full` capture into the per-kernel figures bench.py reports and the judge reads:

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > /tmp/raw.csv
    python tools/ncu_traffic.py /tmp/raw.csv BATCH "source note" [profiles/ncu_traffic.json] > profiles/rNN_ncu_summary.md

Writes profiles/ncu_traffic.json = {kernel key: {"bytes": dram__bytes_read.sum + dram__bytes_write.sum per launch,
"batch": BATCH, "source": ...}} (read by bench.py for roofline.traffic) and prints a markdown summary of the counters
that matter for these kernels (integer-pipe utilisation, issue slots, registers, L2 hit rate, DRAM bytes)."""
import csv
import json
import os
import sys

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
KEYS = {"br_cggi32_kernel": "br_cggi32_kernel", "br_dm32_kernel": "br_dm32_kernel", "br_cggi64w_kernel": "br_cggi64w_kernel",
        "br_cggi64_kernel": "br_cggi64_kernel", "br_generic_kernel": "br_generic_kernel", "mkmswitch": "mkmswitch"}
SHOW = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.avg.per_cycle_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]


def main():
    raw, batch, note = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    out_json = sys.argv[4] if len(sys.argv) > 4 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                                  "profiles", "ncu_traffic.json")
    rows = list(csv.reader(open(raw)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units, data = rows[hdr], rows[hdr + 1], rows[hdr + 2:]
    col = {n: i for i, n in enumerate(names)}
    try:
        traffic = json.load(open(out_json))
    except Exception:
        traffic = {}
    print(f"# ncu --set full summary ({note})\n")
    for r in data:
        if len(r) < len(names):
            continue
        kname = r[col["Kernel Name"]]

        def val(metric):
            if metric not in col:
                return None
            x, u = r[col[metric]].replace(",", ""), units[col[metric]]
            try:
                return float(x) * UNIT.get(u, 1)
            except ValueError:
                return None

        rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
        print(f"## {kname[:110]}\n")
        print("| metric | value | unit |\n|---|---|---|")
        for m in SHOW:
            if m in col:
                print(f"| {m} | {r[col[m]]} | {units[col[m]]} |")
        print()
        for frag, key in KEYS.items():
            if frag in kname and rd is not None and wr is not None and rd == rd and wr == wr:   # NaN: incomplete capture
                traffic[key] = {"bytes": rd + wr, "batch": batch, "kernel": kname[:120], "source": note}
                break
    json.dump(traffic, open(out_json, "w"), indent=1)
    print(f"\nwritten: {out_json}")


if __name__ == "__main__":
    main()
