set -x
python microbench/one.py 60 slot_bits=16 > gpurun_out/one60.log 2>&1 && ncu --set full --clock-control none --import-source on --kernel-name regex:"line_filter|line_probe" --launch-skip 4 --launch-count 2 -o gpurun_out/r02_line2 -f python microbench/one.py 60 slot_bits=16 > gpurun_out/ncu_one60.log 2>&1
tail -n 2 gpurun_out/one60.log gpurun_out/ncu_one60.log
