set -x
python microbench/one.py 300 slot_bits=16 2>&1 | tail -n 1
python microbench/one.py 60 slot_bits=16 2>&1 | tail -n 1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct --clock-control none --kernel-name regex:"line_" --launch-skip 14 --launch-count 7 --csv --log-file gpurun_out/r02_line3_metrics300.csv python microbench/one.py 300 slot_bits=16 > gpurun_out/ncu_one300.log 2>&1
cut -d, -f5,13- gpurun_out/r02_line3_metrics300.csv | grep -v "^\"Kernel" | grep "time_dur\|bytes_read\|issue_act" | tail -n 24
