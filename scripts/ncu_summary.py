"""Summarise an .ncu-rep: headline metrics per kernel + the hottest SASS lines by stall samples.
usage: python scripts/ncu_summary.py REP [n_hot_lines]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; nhot = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size"]
ki = hdr.index("Kernel Name")
for r in rows[2:]:
    print("==", r[ki][:90])
    for w in want:
        if w in hdr:
            i = hdr.index(w); print(f"   {w:75s} {r[i]:>18s} {rows[1][i]}")
    for i, h in enumerate(hdr):
        if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
            try:
                if float(r[i]) > 0.25: print(f"   stall {h.split('issue_stalled_')[1].split('_per_issue')[0]:40s} {float(r[i]):.2f}")
            except ValueError: pass
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
cur = None; data = []
for r in rows:
    if r and r[0] == "Kernel Name": cur = r[1][:60]; continue
    if r and r[0] == "Address": h = r; continue
    if cur and len(r) > 6 and r[h.index("Instructions Executed")].isdigit():
        data.append((cur, int(r[h.index("Warp Stall Sampling (All Samples)")]), int(r[h.index("Instructions Executed")]), r[h.index("Source")].strip()[:70], len(data)))
for k in sorted(set(d[0] for d in data)):
    dk = [d for d in data if d[0] == k]
    tot = sum(d[1] for d in dk) or 1
    print("== hot SASS lines of", k, "total samples", tot, "total inst", sum(d[2] for d in dk))
    for d in sorted(dk, key=lambda d: -d[1])[:nhot]:
        print(f"   {100*d[1]/tot:5.1f}%  exec {d[2]:>10d}  #{d[4]:<5d} {d[3]}")
