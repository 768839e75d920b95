# three-pass line kernels: parity, then slice-size sweep (one process per setting: the size is read once)
set -x
timeout 900 python -m pytest tests/test_gpu_line.py -x -q 2>&1 | tail -3
python microbench/sweep.py --genomes 300 --configs "slot_bits=32" 2>&1 | grep cls=
for cfg in ${CFGS:-1024 2048 4096 8192 16384 32768 100000000}; do
  echo "== SLICE=$cfg"
  KA_LINE_SLICE=$cfg python microbench/sweep.py --genomes 300 --configs "slot_bits=16" 2>&1 | grep cls=
done
