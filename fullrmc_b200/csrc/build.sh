#!/bin/bash
# Builds libfullrmc_b200.so for sm_100a in-tree (fullrmc_b200/lib/).
# -fmad=false: no FMA contraction anywhere (the distance/bin arithmetic must match the
# reference's fp32 operation order bit for bit); IEEE sqrt/div are nvcc defaults and are
# spelled out so that nobody "optimises" them away.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../lib"
mkdir -p "$OUT"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo
       -fmad=false -prec-div=true -prec-sqrt=true -ftz=false
       -Xcompiler -fPIC -Xcompiler -O2 -Xcompiler -fno-fast-math)
if [ "${FRMC_PTXAS_V:-0}" = "1" ]; then FLAGS+=(-Xptxas -v); fi
if [ -n "${FRMC_EXTRA_FLAGS:-}" ]; then FLAGS+=(${FRMC_EXTRA_FLAGS}); fi   # experiments (e.g. -DFRMC_X=1)
OUT="${FRMC_OUT_DIR:-$OUT}"; mkdir -p "$OUT"
OBJS=()
PIDS=()
for f in common stateless fullhist devlayout multigpu store atomdist coordnum storedist storecoord; do
  rm -f "$OUT/$f.o"                     # a failed compile must never leave a stale object for the link step
  "$NVCC" "${FLAGS[@]}" -c "$HERE/$f.cu" -o "$OUT/$f.o" &
  PIDS+=("$!")
  OBJS+=("$OUT/$f.o")
done
for pid in "${PIDS[@]}"; do
  wait "$pid" || { echo "nvcc failed (pid $pid)" >&2; exit 1; }
done
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT/libfullrmc_b200.so" "${OBJS[@]}" -lcudart -ldl
echo "built $OUT/libfullrmc_b200.so"
