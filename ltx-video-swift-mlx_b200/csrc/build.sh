#!/bin/bash
# Builds libltxcuda.so in-tree for sm_100a (nvcc cross-compiles without a GPU).  Usage: build.sh [-j N]
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../libltxcuda.so"
OBJ="$HERE/build"
mkdir -p "$OBJ"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden)
SRCS=(safetensors tmap gemm gemm2 gemm_swapab gemm_skinny gemm_q attention elementwise conv3d weights dit dit_av dit_f32 vae vae_extra dist api)
pids=()
for s in "${SRCS[@]}"; do
  if [ ! -f "$OBJ/$s.o" ] || [ "$HERE/$s.cu" -nt "$OBJ/$s.o" ] || [ -n "$(find "$HERE" -maxdepth 1 \( -name '*.h' -o -name '*.cuh' \) -newer "$OBJ/$s.o" 2>/dev/null)" ] || [ "$HERE/../../include/ltxcuda.h" -nt "$OBJ/$s.o" ]; then
    "$NVCC" "${FLAGS[@]}" -c "$HERE/$s.cu" -o "$OBJ/$s.o" &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
objs=()
for s in "${SRCS[@]}"; do objs+=("$OBJ/$s.o"); done
"$NVCC" -shared -gencode arch=compute_100a,code=sm_100a -o "$OUT" "${objs[@]}" -Xcompiler -fvisibility=hidden -ldl
echo "built $OUT"
