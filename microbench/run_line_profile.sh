# ncu evidence of the three line-table passes (60 proteomes, third annotate) + the miss-heavy comparison
set -x
python microbench/one.py 60 slot_bits=16 > gpurun_out/one60.log 2>&1 && ncu --set full --clock-control none --import-source on --kernel-name regex:"line_filter|line_probe|line_tally" --launch-skip 8 --launch-count 3 -o gpurun_out/r02_line_passes -f python microbench/one.py 60 slot_bits=16 > gpurun_out/ncu_one60.log 2>&1
tail -n 2 gpurun_out/one60.log gpurun_out/ncu_one60.log
ncu -i gpurun_out/r02_line_passes.ncu-rep --page raw --csv > gpurun_out/r02_line_passes_raw.csv
( for sb in 32 16; do for seed in 0 7; do python microbench/one.py 300 slot_bits=$sb batch_seed=$seed | tail -n 1; done; done ) > gpurun_out/r02_line_vs_sector.log 2>&1
cat gpurun_out/r02_line_vs_sector.log
