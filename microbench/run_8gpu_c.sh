# third 8-GPU session: line table as the default layout, segment-coalesced routed scatter
set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 8 --warmup 3 > gpurun_out/bench_n8_c.json 2> gpurun_out/bench_n8_c.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_n8_c.json").read().strip().split("\n")[-1])
print("value", d["value"]/1e6, d["ms_per_step"], "e2e", d["e2e"]["value"]/1e6, d["e2e"]["ms_per_step"], "e2e_bytes", d["e2e_bytes"]["value"]/1e6, d["e2e_bytes"]["ms_per_step"])
print("ingest", d.get("ingest_paths"))
print("multi", {k:(v["e2e_ms"], v["matches_single_gpu_results"]) for k,v in d["multi_device_engine"].items() if isinstance(v,dict)})
print("c5small", d["c5_small_parity"])
for m in (1,2,3):
    x=d["c5_sharded"][f"table_mode_{m}"]; print("c5 mode",m, x["e2e_ms"], x["probes_per_s"]/1e9, x["e2e_bytes_ms"], x["byte_form_identical"])
print(d["c5_sharded"]["modes_identical_on_all_proteins"])
PY
tail -n 3 gpurun_out/bench_n8_c.err
python microbench/c5_oversized.py 8 1.2e10 4000000 3 > gpurun_out/r02_c5_oversized_8gpu_c.log 2>&1; grep -v NCCL gpurun_out/r02_c5_oversized_8gpu_c.log | cut -c1-330
