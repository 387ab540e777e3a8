# Timing-only ablations of the fused MLP kernel (NVP_ABL bits, see mlp_fused.cuh).
#   bash scripts/fused_ablate.sh build     (here: builds abl_build/libnvp_b200_ablN.so for every variant)
#   bash scripts/fused_ablate.sh run TAG   (through gpurun: per-kind times of every variant -> gpurun_out/TAG_ablate.log)
VARIANTS=${VARIANTS:-"0 1 2 4 8 16 32 3 7 15 63"}
NVCC=/usr/local/cuda/bin/nvcc
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr"
if [ "$1" = build ]; then
  mkdir -p build/abl abl_build
  for v in $VARIANTS; do
    ( $NVCC $FLAGS -DNVP_ABL=$v -c nvp_b200/csrc/mlp_tc.cu -o build/abl/mlp_tc_$v.o &&
      $NVCC -gencode arch=compute_100a,code=sm_100a -shared $(ls build/obj/*.o | grep -v mlp_tc.o) build/abl/mlp_tc_$v.o -o abl_build/libnvp_b200_abl$v.so -lcuda ) &
  done
  ( $NVCC $FLAGS -DNVP_FCONST=1 -c nvp_b200/csrc/mlp_tc.cu -o build/abl/mlp_tc_ldc.o &&
    $NVCC -gencode arch=compute_100a,code=sm_100a -shared $(ls build/obj/*.o | grep -v mlp_tc.o) build/abl/mlp_tc_ldc.o -o abl_build/libnvp_b200_ablldc.so -lcuda ) &
  wait
  ls -la build/abl/*.so
else
  TAG=${2:-abl}
  mkdir -p gpurun_out
  : > gpurun_out/${TAG}_ablate.log
  for v in $VARIANTS ldc; do
    echo "== NVP_ABL=$v" >> gpurun_out/${TAG}_ablate.log
    NVP_B200_LIB=$PWD/abl_build/libnvp_b200_abl$v.so timeout -k 5 120 python scripts/grid_sweep.py >> gpurun_out/${TAG}_ablate.log 2>&1
  done
  cat gpurun_out/${TAG}_ablate.log
fi
