set -x
# config-4 shape (single breakend: variant templates of 601 rows -> R = 19, reference templates of 400 rows -> R = 13): the R = 19 launch
# (selected by its mangled name; time_shape.py was run without ncu first)
timeout -s KILL 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:sw_wavefront_kernelILi19E --launch-count 1 -o gpurun_out/r02f_fwd_sbnd_r19 python tools/time_shape.py single_breakend 600 > gpurun_out/r02f_ncu_sbnd.log 2>&1
tail -3 gpurun_out/r02f_ncu_sbnd.log
