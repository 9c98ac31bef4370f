#!/usr/bin/env bash
# Builds the product libraries IN-TREE for sm_100a (the .so files travel to the GPU box with the snapshot).
#   ekf-slam-ml_b200/libekfslam_b200.so          C ABI of include/ekf_slam_b200.h + include/circle_fit_b200.h
#   ekf-slam-ml_b200/libekfslam_sharded_b200.so  C ABI of include/ekf_sharded_b200.h (row-sharded filter; links NCCL)
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
SRC=ekf-slam-ml_b200/csrc
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fno-gnu-unique -Xlinker --version-script=$SRC/exports.map"
newest=$(ls -t $SRC/*.cu $SRC/*.cuh $SRC/sharded/*.cu $SRC/exports.map include/*.h build.sh | head -1)

OUT=ekf-slam-ml_b200/libekfslam_b200.so
if [ -f "$OUT" ] && [ "$OUT" -nt "$newest" ] && [ "${FORCE:-0}" != "1" ]; then
  echo "up to date: $OUT"
else
  $NVCC $FLAGS ${EXTRA_NVCC_FLAGS:-} -shared -o $OUT $(ls $SRC/*.cu) -lcudart
  echo "built $OUT"
fi

OUT2=ekf-slam-ml_b200/libekfslam_sharded_b200.so
if [ -f "$OUT2" ] && [ "$OUT2" -nt "$newest" ] && [ "${FORCE:-0}" != "1" ]; then
  echo "up to date: $OUT2"
else
  # libnccl.so.2 is resolved at load time: torch's bundled NCCL when torch is already imported, the system one otherwise
  $NVCC $FLAGS ${EXTRA_NVCC_FLAGS:-} -shared -o $OUT2 $SRC/sharded/sharded.cu -lcudart -lnccl
  echo "built $OUT2"
fi
