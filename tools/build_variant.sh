#!/usr/bin/env bash
# Builds an experimental variant of libpe_b200.so with extra -D flags into physicedit_b200/lib/variants/<name>.so
# (select it at run time with PE_B200_LIB=<path>).  Usage: tools/build_variant.sh <name> -DFOO=1 ...
set -euo pipefail
root="$(cd "$(dirname "$0")/.." && pwd)"
name="$1"; shift
out="$root/physicedit_b200/lib/variants"; mkdir -p "$out/obj_$name"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -cudart shared)
pids=()
for f in capi rowwise vae_kernels llm_kernels train_kernels gemm_sm100 attention_sm100; do
  /usr/local/cuda/bin/nvcc "${FLAGS[@]}" "$@" -c "$root/physicedit_b200/csrc/$f.cu" -o "$out/obj_$name/$f.o" 2>/dev/null &
  pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
/usr/local/cuda/bin/nvcc -shared -cudart shared -o "$out/$name.so" "$out/obj_$name"/*.o
rm -rf "$out/obj_$name"
echo "$out/$name.so"
