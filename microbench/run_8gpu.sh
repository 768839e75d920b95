# one 8-GPU session: ingest ceiling, sharded-table parity, config 5 at full size, bench.py at N = 8
set -x
nvidia-smi topo -m > gpurun_out/r02_topo.txt 2>&1
python microbench/pcie_concurrent.py > gpurun_out/r02_pcie_concurrent.log 2>&1; cat gpurun_out/r02_pcie_concurrent.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cli.py tests/test_gpu_distance.py -x -q -k "routed or sharded or multi_device" > gpurun_out/r02_pytest_sharded_8gpu.log 2>&1; tail -3 gpurun_out/r02_pytest_sharded_8gpu.log
python microbench/c5_oversized.py 8 1.2e10 4000000 1,2,3 > gpurun_out/r02_c5_oversized_8gpu.log 2>&1; cat gpurun_out/r02_c5_oversized_8gpu.log
KA_C5_CHUNK=33554432 python microbench/c5_oversized.py 8 1.2e10 4000000 3 > gpurun_out/r02_c5_oversized_8gpu_chunk32.log 2>&1; tail -2 gpurun_out/r02_c5_oversized_8gpu_chunk32.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 8 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
tail -c 3000 gpurun_out/bench_n8.json; tail -3 gpurun_out/bench_n8.err
