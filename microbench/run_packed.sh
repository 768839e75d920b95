set -x
timeout 900 python -m pytest tests/test_gpu_line.py tests/test_gpu_parity.py tests/test_gpu_hypothesis.py -x -q 2>&1 | tail -6
python microbench/sweep.py --genomes 300 --configs "slot_bits=32;slot_bits=32,resident_packed=0" 2>&1 | tee gpurun_out/sweep_packed.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().split('\n')[-1])
print('value', d['value']/1e6, 'ms', d['ms_per_step'], 'e2e', d['e2e']['value']/1e6, d['e2e']['ms_per_step'], 'e2e_bytes', d['e2e_bytes']['value']/1e6, d['e2e_bytes']['ms_per_step'], 'pack', d['host_pack'], 'c2', d['c2_single_proteome'])
PY
tail -3 gpurun_out/bench_n1.err
