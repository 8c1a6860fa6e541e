"""Summarise an .ncu-rep (raw page) into the few metrics DESIGN.md/profiles cite: python scripts/ncu_extract.py rep"""
import csv, subprocess, sys
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_elapsed.max",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "sm__warps_active.avg.pct_of_peak_sustained_active"]
idx = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    print("----")
    for w in want:
        if w in idx:
            print(f"{w:72s} {r[idx[w]]:>22s} {units[idx[w]]}")
