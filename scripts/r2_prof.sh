# usage (through gpurun): bash scripts/r2_prof.sh TAG   -> launch lists + ncu --set full captures for configs S and L
TAG=${1:-r2prof}
mkdir -p gpurun_out
for cfg in s l; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${TAG}_launches_${cfg}.csv python scripts/prof_step.py 3 $cfg > gpurun_out/${TAG}_prof_${cfg}.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:"grid_|mlp_|pack_|fused_consts" -s 10 -c 12 -f -o gpurun_out/${TAG}_full_${cfg} python scripts/prof_step.py 2 $cfg > gpurun_out/${TAG}_ncu_${cfg}.log 2>&1
  tail -2 gpurun_out/${TAG}_ncu_${cfg}.log
done
