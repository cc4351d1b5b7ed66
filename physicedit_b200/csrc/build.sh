#!/usr/bin/env bash
# Builds libpe_b200.so (sm_100a only) in-tree.  Usage: build.sh [extra nvcc flags]
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"
out="$here/../lib"
mkdir -p "$out"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-Wall
       --expt-relaxed-constexpr -Xptxas -v -cudart shared)
objs=()
pids=()
for f in capi rowwise vae_kernels llm_kernels train_kernels gemm_sm100 attention_sm100; do
  "$NVCC" "${FLAGS[@]}" "$@" -c "$here/$f.cu" -o "$out/$f.o" > "$out/$f.ptxas.log" 2>&1 &
  pids+=($!)
  objs+=("$out/$f.o")
done
rc=0
for p in "${pids[@]}"; do wait "$p" || rc=1; done
if [ $rc -ne 0 ]; then cat "$out"/*.ptxas.log | grep -v '^ptxas info' >&2 || true; exit 1; fi
"$NVCC" -shared -cudart shared -o "$out/libpe_b200.so" "${objs[@]}"
echo "built $out/libpe_b200.so"
