# usage: bash scripts/gpu_ncu.sh TAG KERNEL_REGEX [skip] [count]  -> gpurun_out/TAG.ncu-rep (full set, with source)
TAG=$1; RE=$2; SKIP=${3:-1}; CNT=${4:-1}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$RE -s $SKIP -c $CNT -f -o gpurun_out/$TAG python scripts/prof_step.py 2 > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log
