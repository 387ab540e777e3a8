# usage (through gpurun): bash scripts/r2_run.sh TAG   -> gpurun_out/TAG_*  (short guarded stages: a hung kernel only costs its own timeout)
TAG=${1:-r2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.log 2>&1
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/sin_probe scripts/sin_probe.cu && timeout -k 5 60 /tmp/sin_probe > gpurun_out/${TAG}_sin_probe.log 2>&1
timeout -k 5 120 python -m pytest tests/test_gpu_parity.py -x -q -k "fused_step and tc and s_" > gpurun_out/${TAG}_pytest_first.log 2>&1
echo "first rc=$?" >> gpurun_out/${TAG}_pytest_first.log
tail -5 gpurun_out/${TAG}_pytest_first.log
timeout -k 5 600 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
timeout -k 5 200 python scripts/grid_sweep.py > gpurun_out/${TAG}_kinds_fused.log 2>&1
NVP_MLP_FUSED=0 timeout -k 5 200 python scripts/grid_sweep.py > gpurun_out/${TAG}_kinds_3k.log 2>&1
cat gpurun_out/${TAG}_kinds_fused.log gpurun_out/${TAG}_kinds_3k.log
timeout -k 5 400 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 1500 gpurun_out/${TAG}_bench.json
