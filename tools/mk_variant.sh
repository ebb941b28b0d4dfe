#!/bin/bash
# usage: tools/mk_variant.sh name [extra nvcc flags]   -> build/variants/lib_<name>.so + fast-path instruction mix
n=$1; shift
root="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
mkdir -p "$root/build/variants"; cd "$root/nanomonsv_b200/csrc" || exit 1
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared "$@" -Xptxas -v -o "$root/build/variants/lib_$n.so" nmsv_api.cu 2>&1 | grep -E "error|sw_wavefront_kernelILi16ELi3ELi1" -A1 | grep -E "error|registers"
python "$root/tools/sass_hot_loop.py" "$root/build/variants/lib_$n.so"
