# usage (through gpurun): bash scripts/r2_full.sh TAG   -> full GPU test suite + bench (S and L) -> gpurun_out/TAG_*
TAG=${1:-r2}
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
timeout -k 5 300 python bench.py --steps 50 --warmup 5 > gpurun_out/${TAG}_bench_s.json 2> gpurun_out/${TAG}_bench_s.err
timeout -k 5 300 python bench.py --steps 30 --warmup 5 --config l --no-cpu-baseline > gpurun_out/${TAG}_bench_l.json 2> gpurun_out/${TAG}_bench_l.err
python scripts/show_bench.py gpurun_out/${TAG}_bench_s.json gpurun_out/${TAG}_bench_l.json 2>&1 | tail -30
