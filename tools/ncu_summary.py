"""Markdown table of the headline counters of one or more .ncu-rep files (read on the CPU box)."""
import csv, subprocess, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.per_cycle_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:  # one row per captured launch
        if len(vals) < len(hdr):
            continue
        d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
        print(f"## {d.get('Kernel Name', ('?',''))[0][:90]}  (`{rep}`)\n\n| metric | value | unit |\n|---|---:|---|")
        for k in KEYS:
            if k in d:
                print(f"| `{k}` | {d[k][0]} | {d[k][1]} |")
        try:
            tr = float(d["dram__bytes_read.sum"][0].replace(",", "")) + float(d["dram__bytes_write.sum"][0].replace(",", ""))
            print(f"| **traffic = dram read + write** | {tr:.1f} | {d['dram__bytes_read.sum'][1]} |")
        except Exception:
            pass
        print()
