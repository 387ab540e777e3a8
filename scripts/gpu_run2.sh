set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r1b_launches.csv python prof_step.py 3 > gpurun_out/r1b_prof.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:grid_binned -s 2 -c 2 -o gpurun_out/r1b_binned python prof_step.py 2 >> gpurun_out/r1b_prof.log 2>&1
tail -2 gpurun_out/r1b_prof.log
