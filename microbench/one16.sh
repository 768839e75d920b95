python microbench/one.py 60 slot_bits=16 > gpurun_out/one60.log 2>&1 && ncu --set full --clock-control none --import-source on --kernel-name regex:line_tile --launch-skip 2 --launch-count 1 -o gpurun_out/r02_line_tile -f python microbench/one.py 60 slot_bits=16 > gpurun_out/ncu_one60.log 2>&1
tail -n 2 gpurun_out/one60.log
