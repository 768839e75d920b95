# two-pass line kernels: parity, sweep against the sector kernel, per-kernel times
set -x
timeout 900 python -m pytest tests/test_gpu_line.py -x -q 2>&1 | tail -15
python microbench/sweep.py --genomes 300 --configs "${SWEEP:-slot_bits=32;slot_bits=16;slot_bits=16,filter=0;slot_bits=16,tile_span=2048,long_seq=2048;slot_bits=16,tile_span=3072,long_seq=3072}" > gpurun_out/sweep_line2.log 2>&1
cat gpurun_out/sweep_line2.log
python microbench/one.py 60 slot_bits=16 > gpurun_out/one60.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,sm__inst_executed_pipe_lsu.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct --clock-control none --kernel-name regex:"line_filter|line_probe|line_tally" --launch-skip 6 --launch-count 3 --csv --log-file gpurun_out/r02_line2_metrics.csv python microbench/one.py 60 slot_bits=16 > gpurun_out/ncu_one60.log 2>&1
tail -n 2 gpurun_out/one60.log
cat gpurun_out/r02_line2_metrics.csv | cut -d, -f5,13- | tail -n 20
