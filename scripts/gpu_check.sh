# usage: bash scripts/gpu_check.sh TAG [pytest-filter]   (run through gpurun; writes gpurun_out/TAG_*)
TAG=${1:-chk}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q ${2:+-k "$2"} 2>&1 | tail -15 > gpurun_out/${TAG}_pytest.log
timeout 400 python bench.py --steps 50 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv python scripts/prof_step.py 3 > gpurun_out/${TAG}_prof.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log
