# line-table parity, then the option sweep on the C3 batch (300 proteomes) and one ncu capture
set -x
timeout 600 python -m pytest tests/test_gpu_line.py -x -q 2>&1 | tail -15
python microbench/sweep.py --genomes 300 --configs "${SWEEP:-;filter=0;slot_bits=32;tile_span=1024,long_seq=1024;tile_span=1024,long_seq=2048;tile_span=2048,long_seq=2048}" > gpurun_out/sweep_line.log 2>&1
cat gpurun_out/sweep_line.log
if [ -n "$LIBB" ]; then KMERANNO_LIB=$PWD/kmers.anno_b200/libkmeranno_b.so python microbench/sweep.py --genomes 300 --configs ";tile_span=1024,long_seq=1024" 2>&1 | tee gpurun_out/sweep_line_b.log; fi
if [ -n "$NCU" ]; then
python microbench/one.py 60 > gpurun_out/one60.log 2>&1 && ncu --set full --clock-control none --import-source on --kernel-name regex:line_tile --launch-skip 2 --launch-count 1 -o gpurun_out/r02_line_tile -f python microbench/one.py 60 > gpurun_out/ncu_one60.log 2>&1
tail -n 3 gpurun_out/one60.log gpurun_out/ncu_one60.log
fi
