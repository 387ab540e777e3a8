set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r1b_pytest.log
timeout 400 python bench.py --steps 50 --warmup 5 > gpurun_out/r1b_bench_tb64.json 2> gpurun_out/r1b_bench_tb64.err
NVP_BIN_TB=128 timeout 400 python bench.py --steps 50 --warmup 5 > gpurun_out/r1b_bench_tb128.json 2> gpurun_out/r1b_bench_tb128.err
tail -3 gpurun_out/r1b_pytest.log
