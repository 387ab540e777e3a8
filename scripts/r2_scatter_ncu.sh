# usage (through gpurun): bash scripts/r2_scatter_ncu.sh TAG -> ncu --set full of the scatter-add kernel, packed and interleaved windows
TAG=${1:-sc}
mkdir -p gpurun_out
for ilv in 0 1; do
  NVP_BIN_ILV=$ilv timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:grid_binned_kernel -s 3 -c 1 -f \
    -o gpurun_out/${TAG}_scatter_ilv${ilv} python scripts/prof_step.py 2 s > gpurun_out/${TAG}_scatter_ilv${ilv}.log 2>&1
  tail -1 gpurun_out/${TAG}_scatter_ilv${ilv}.log
done
