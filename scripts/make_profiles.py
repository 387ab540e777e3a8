"""Turn one session's gpurun_out/ evidence into the committed profiles/ files.

usage: python scripts/make_profiles.py TAG NCU_REP LAUNCHES_CSV [CONFIG]
  NCU_REP       `ncu --set full --clock-control none --import-source on` capture of one training step (scripts/gpu_ncu.sh)
  LAUNCHES_CSV  `ncu --metrics gpu__time_duration.sum --clock-control none` launch list of prof_step.py (scripts/gpu_check.sh)
writes profiles/TAG_ncu_full_summary.csv, profiles/TAG_launches.csv, profiles/TAG_traffic_CONFIG.json
"""
import csv, io, json, os, shutil, subprocess, sys

tag, rep, launches = sys.argv[1:4]
config = sys.argv[4] if len(sys.argv) > 4 else "s"
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
keep = ["ID", "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sectors.sum"]
idx = [hdr.index(k) for k in keep if k in hdr]
with open(os.path.join(root, f"{tag}_ncu_full_summary.csv"), "w", newline="") as f:
    w = csv.writer(f)
    for r in rows:
        w.writerow([r[i] for i in idx])
shutil.copyfile(launches, os.path.join(root, f"{tag}_launches.csv"))

kind_of = {"grid_binned_kernel<2, 0": "grid_gather", "grid_binned_kernel<4, 0": "grid_gather", "grid_binned_kernel<2, 1": "grid_scatter",
           "grid_binned_kernel<4, 1": "grid_scatter", "gather_": "grid_gather", "grid_gather": "grid_gather", "scatter": "grid_scatter",
           "mlp_fused": "mlp_fused", "mlp_forward": "mlp_forward", "mlp_backward": "mlp_backward", "mlp_wgrad": "mlp_wgrad", "grid_bin_": "grid_bin",
           "pack_weights": "pack"}
ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
unit = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
traffic = {}
for r in rows[2:]:
    kind = next((v for k, v in kind_of.items() if k in r[ki]), None)
    if kind is None:
        continue
    b = float(r[ri]) * unit[rows[1][ri]] + float(r[wi]) * unit[rows[1][wi]]
    traffic[kind] = traffic.get(kind, 0.0) + b
with open(os.path.join(root, f"{tag}_traffic_{config}.json"), "w") as f:
    json.dump({"source": f"ncu --set full --clock-control none capture of `python scripts/prof_step.py 2` (one step's kernels), {tag}; "
                         "dram__bytes_read.sum + dram__bytes_write.sum summed over the kernels of each kind",
               "dram_bytes_per_step": traffic}, f, indent=1)
print(json.dumps(traffic, indent=1))
