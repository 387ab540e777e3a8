# usage (through gpurun): bash scripts/r2_quick.sh TAG  -> parity tests of the tensor-core path + per-kind times
TAG=${1:-r2q}
mkdir -p gpurun_out
timeout -k 5 150 python -m pytest tests/test_gpu_parity.py -x -q -k "tc" > gpurun_out/${TAG}_pytest_tc.log 2>&1
echo "rc=$?" >> gpurun_out/${TAG}_pytest_tc.log
tail -4 gpurun_out/${TAG}_pytest_tc.log
timeout -k 5 200 python scripts/grid_sweep.py > gpurun_out/${TAG}_kinds.log 2>&1
cat gpurun_out/${TAG}_kinds.log
