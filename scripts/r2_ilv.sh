# usage (through gpurun): bash scripts/r2_ilv.sh TAG  -> scatter-add layout A/B: grid parity tests, then bench kernel times per variant
TAG=${1:-ilv}
mkdir -p gpurun_out
[ -n "$SKIP_TESTS" ] || timeout -k 5 600 python -m pytest tests/test_gpu_binned.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
run() {  # name, env...
  name=$1; shift
  env "$@" timeout -k 5 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_${name}.json 2> gpurun_out/${TAG}_${name}.err
  python - "$name" gpurun_out/${TAG}_${name}.json <<'P'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    k = {n: x["ms_per_step"] for n, x in d["kernels"].items()}
    print(f'{sys.argv[1]:>14}: step {d["ms_per_step"]:.3f} ms  scatter {k.get("grid_scatter", 0):.4f}  gather {k.get("grid_gather", 0):.4f}  bin {k.get("grid_bin", 0):.4f}')
except Exception as e:
    print(sys.argv[1], "failed", e)
P
}
run ilv16_u4 NVP_BIN_ILV=1
run ilv16_u8 NVP_B200_LIB=$PWD/abl_build/libnvp_b200_u8.so
run ilv16_u8_sp1 NVP_B200_LIB=$PWD/abl_build/libnvp_b200_u8.so NVP_BIN_SPARSE_WARPS_S=1
run ilv15_u8_sp1 NVP_B200_LIB=$PWD/abl_build/libnvp_b200_u8.so NVP_BIN_SPARSE_WARPS_S=1 NVP_BIN_WARPS_ILV=15
run packed22_u8_sp1 NVP_B200_LIB=$PWD/abl_build/libnvp_b200_u8.so NVP_BIN_ILV=0 NVP_BIN_SPARSE_WARPS_S=1
