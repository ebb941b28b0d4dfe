# usage (on a GPU box): bash tools/run_variants.sh v1 v2 ...   -> gpurun_out/variants.log  (libraries from tools/mk_variant.sh)
: > gpurun_out/variants.log
for v in "$@"; do NMSV_LIB_PATH=build/variants/lib_$v.so timeout 300 python tools/time_kernel.py ${NMSV_TIME_N:-384} >> gpurun_out/variants.log 2>&1; done
grep fwd_ms gpurun_out/variants.log
