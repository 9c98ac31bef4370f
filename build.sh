#!/usr/bin/env bash
# Builds the product library IN-TREE for sm_100a (the .so travels to the GPU box with the snapshot).
#   ekf-slam-ml_b200/libekfslam_b200.so   C ABI of include/ekf_slam_b200.h (+ circle fitting)
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
SRC=ekf-slam-ml_b200/csrc
OUT=ekf-slam-ml_b200/libekfslam_b200.so
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=default"
SOURCES=$(ls $SRC/*.cu)
newest=$(ls -t $SRC/*.cu $SRC/*.cuh include/*.h build.sh | head -1)
if [ -f "$OUT" ] && [ "$OUT" -nt "$newest" ] && [ "${FORCE:-0}" != "1" ]; then
  echo "up to date: $OUT"; exit 0
fi
$NVCC $FLAGS ${EXTRA_NVCC_FLAGS:-} -shared -o $OUT $SOURCES -lcudart
echo "built $OUT"
