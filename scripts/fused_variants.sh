# Build and time variants of the tensor-core MLP kernels (extra -D flags for mlp_tc.cu), e.g.
#   bash scripts/fused_variants.sh build "base= gt=-DNVP_FOPT=1 p3h=-DNVP_FOPT=2"      (here)
#   bash scripts/fused_variants.sh run TAG "base gt p3h"                                (through gpurun)
NVCC=/usr/local/cuda/bin/nvcc
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr"
if [ "$1" = build ]; then
  mkdir -p build/abl abl_build
  for spec in $2; do
    name=${spec%%=*}; defs=${spec#*=}
    ( $NVCC $FLAGS $(echo $defs | tr ',' ' ') -c nvp_b200/csrc/mlp_tc.cu -o build/abl/mlp_tc_$name.o &&
      $NVCC -gencode arch=compute_100a,code=sm_100a -shared $(ls build/obj/*.o | grep -v mlp_tc.o) build/abl/mlp_tc_$name.o -o abl_build/libnvp_b200_$name.so -lcuda ) &
  done
  wait
  ls abl_build/
else
  TAG=$2
  mkdir -p gpurun_out
  : > gpurun_out/${TAG}_variants.log
  for name in $3; do
    echo "== $name" >> gpurun_out/${TAG}_variants.log
    NVP_B200_LIB=$PWD/abl_build/libnvp_b200_$name.so timeout -k 5 120 python scripts/grid_sweep.py >> gpurun_out/${TAG}_variants.log 2>&1
    NVP_B200_LIB=$PWD/abl_build/libnvp_b200_$name.so timeout -k 5 120 python -m pytest tests/test_gpu_parity.py -x -q -k "fused_step and tc and s_" 2>&1 | tail -1 >> gpurun_out/${TAG}_variants.log
  done
  cat gpurun_out/${TAG}_variants.log
fi
