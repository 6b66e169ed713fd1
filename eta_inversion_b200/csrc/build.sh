#!/usr/bin/env bash
# Builds libetai.so (sm_100a only) next to this script. Usage: build.sh [-j N]
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr -DETAI_BUILD"
SRCS="elementwise textvae_ops backward norm gemm_simt attn_simt cross_attn gemm_tc attn_tc cross_attn_tc unet vae clip api"
mkdir -p build
pids=()
for s in $SRCS; do
  if [ ! -f build/$s.o ] || [ $s.cu -nt build/$s.o ] || [ common.cuh -nt build/$s.o ] || [ ops.cuh -nt build/$s.o ] || [ tc_common.cuh -nt build/$s.o ] || [ engine_base.cuh -nt build/$s.o ] || [ ../../include/etai.h -nt build/$s.o ]; then
    ( $NVCC $FLAGS ${PTXAS_V:+-Xptxas -v} -c $s.cu -o build/$s.o 2> build/$s.log || { cat build/$s.log; exit 1; } ) &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait $p; done
objs=""; for s in $SRCS; do objs="$objs build/$s.o"; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o ../libetai.so $objs -lcudart
echo "built $(realpath ../libetai.so)"
