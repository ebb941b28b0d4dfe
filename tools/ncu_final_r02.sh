set -x
# final round-2 build: the forward launch of the headline batch (m = 2000, default pruning), the sw_jump warp kernel, and the
# launch list of the bench command -- each only after the same command has exited 0 without ncu
timeout -s KILL 300 python tools/time_kernel.py 384 > gpurun_out/r02f_time_kernel.log 2>&1 || exit 1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:sw_wavefront_kernel --launch-count 1 -o gpurun_out/r02f_fwd_m2000 python tools/time_kernel.py 384 > gpurun_out/r02f_ncu_m2000.log 2>&1
timeout -s KILL 300 python tools/time_jump.py > gpurun_out/r02f_time_jump.log 2>&1 || exit 1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:sw_jump_warp_kernel --launch-skip 2 --launch-count 1 -o gpurun_out/r02f_sw_jump python tools/time_jump.py > gpurun_out/r02f_ncu_jump.log 2>&1
timeout -s KILL 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r02f_bench_for_launches.json 2> gpurun_out/r02f_bench_for_launches.err || exit 1
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02f_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r02f_ncu_launches.log 2>&1
ls -la gpurun_out/r02f_*
