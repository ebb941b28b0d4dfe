set -x
# m = 2000 (headline batch, default pruning), m = 400 (config 1 shape), m = 6000 (config 3 shape)
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:sw_wavefront_kernel --launch-count 1 -o gpurun_out/r02_fwd_m2000 python tools/time_kernel.py 384 > gpurun_out/r02_ncu_m2000.log 2>&1
NMSV_BC_ONLY=testdata_shaped timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:sw_wavefront_kernel --launch-count 1 -o gpurun_out/r02_fwd_m400 python tools/time_shape.py testdata_shaped 300 > gpurun_out/r02_ncu_m400.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:sw_wavefront_kernel --launch-count 1 -o gpurun_out/r02_fwd_m6000 python tools/time_shape.py insertion_heavy 90 > gpurun_out/r02_ncu_m6000.log 2>&1
# launch list of the bench command
timeout -s KILL 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r02_bench_for_launches.json 2> gpurun_out/r02_bench_for_launches.err
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r02_ncu_launches.log 2>&1
ls -la gpurun_out/r02_*
