# ncu evidence of the sector class 32 tile kernel (slot_bits = 32) for profiles/
set -x
python microbench/one.py 60 slot_bits=32 > gpurun_out/one60_tile.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on --kernel-name regex:tile_kernel --launch-skip 2 --launch-count 1 -o gpurun_out/r02_tile_kernel -f python microbench/one.py 60 slot_bits=32 > gpurun_out/ncu_tile.log 2>&1
ncu -i gpurun_out/r02_tile_kernel.ncu-rep --page raw --csv > gpurun_out/r02_tile_kernel_raw.csv 2>/dev/null
tail -n 2 gpurun_out/one60_tile.log gpurun_out/ncu_tile.log
