# usage (through gpurun): bash scripts/r2_scatter_split.sh TAG -> durations of the window role and the voxel role of the scatter-add as separate kernels
TAG=${1:-split}
mkdir -p gpurun_out
for ilv in 0 1; do
  NVP_BIN_ILV=$ilv NVP_BIN_SPARSE_WARPS_S=0 timeout -k 5 200 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum --clock-control none -k regex:"grid_binned|sparse_scatter" -c 8 --csv \
    --log-file gpurun_out/${TAG}_ilv${ilv}.csv python scripts/prof_step.py 2 s > gpurun_out/${TAG}_ilv${ilv}.log 2>&1
  grep -v "^==" gpurun_out/${TAG}_ilv${ilv}.csv | cut -d, -f5,13- | tail -16
done
