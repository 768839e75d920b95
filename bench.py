#!/usr/bin/env python
"""bench.py — k-mer annotation hot path on B200: sequences/s and k-mer probes/s.

    python bench.py --gpus N --steps K --warmup W            (our arm; N>1 under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (BASELINE.json configs[2] with the configs[1] table): every GPU annotates, per
step, a batch of `--genomes` synthetic bacterial proteomes (4,500 proteins, ~1.4 M aa each;
SURVEY.md §8d generator, seed 20261018) against a replicated synthetic signature table of
`--table-kmers` (1e8) 8-mers and 30,000 roles.  Weak scaling: the per-GPU batch is fixed,
ranks share nothing on the data path (no collective), results are gathered by the host.

  value / ms_per_step  inputs already resident in HBM: plan + tile (+ long-sequence)
                       kernels per step, timed with CUDA events on the engine's own stream
                       inside libkmeranno.so (ka_get_stats), max over ranks.
  e2e                  the same batch through the public C-ABI call ka_annotate() with
                       PINNED HOST buffers: chunked H2D of residues+offsets, kernels and D2H
                       of the per-sequence results are all inside the timed region.
  roofline             dominant kernel = tile_kernel; achieved = 33 B/probe x probes per
                       launch / its mean CUDA-event duration; peak = MEASURED_PEAKS.json.
  rand_roofline        the graded denominator of BASELINE.md §3: R_rand = independent random
                       32-byte sector loads over a buffer the size of the table, measured
                       live in this run by ka_probe_roofline.
  cpu_baseline         the Java-shaped oracle (oracle/, `port`: no JVM exists here) timed on
                       the host cores on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 20261018
N_PROT = 4500
BYTES_PER_PROBE = 33.0  # 32-byte bucket sector + 1 residue byte
# dram__bytes_read.sum + dram__bytes_write.sum of tile_kernel<32,4,128,6> from the ncu --set full
# capture of `bench.py --genomes 60` (profiles/r01_summary.md §E): 8.426 GB (8.417 read + 0.009 write) for 88.1 M probes.
# B200 fills a whole 128-byte line per L2 miss, hence ~3x the algorithmic bytes.
NCU_DRAM_BYTES_PER_PROBE = 8.426e9 / 88.1e6


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--genomes", type=int, default=1000, help="proteomes per GPU per step")
    ap.add_argument("--table-kmers", type=float, default=1e8)
    ap.add_argument("--roles", type=int, default=30000)
    ap.add_argument("--K", type=int, default=8)
    ap.add_argument("--min-hits", type=int, default=5)
    ap.add_argument("--mode", type=int, default=0, choices=[0, 1, 2],
                    help="0 = C3 proteomes; 1 / 2 = config-4 skewed lengths (log-uniform / bimodal 50..5000 aa)")
    ap.add_argument("--cpu-genomes", type=int, default=384, help="proteomes in the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cpu-best", action="store_true", help="skip the packed-integer 'best CPU' line")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--option", action="append", default=[], help="engine option name=value")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.lines = []
        self.proc = None
        self.first = 0

    def mark(self):
        self.first = len(self.lines)

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for l in self.lines[self.first:]:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v == "Active":
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def bind_to_gpu_cpus(gpu_index):
    """Pin this process to the CPUs NVML reports as local to its GPU so that the pinned host
    buffers (first touch) and the H2D copies stay on the GPU's NUMA node.  Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, 16)
        cpus = {64 * w + b for w, v in enumerate(words) for b in range(64) if (int(v) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if use and use != allowed:
            os.sched_setaffinity(0, use)
            return f"bound to {len(use)} GPU-local CPUs"
        return "no narrower GPU-local CPU set"
    except Exception as e:  # noqa: BLE001
        return f"unbound ({type(e).__name__})"


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def workload_name(a):
    shape = {0: "C3 batch", 1: "C4 batch (log-uniform 50..5000 aa)", 2: "C4 batch (bimodal 50..300 / 3000..5000 aa)"}[a.mode]
    return (f"{shape}: {a.genomes} synthetic proteomes x {N_PROT} proteins per GPU per step vs replicated "
            f"C2 table ({a.table_kmers:.0e} {a.K}-mers, {a.roles} roles)")


def make_table(a):
    from kmers_anno_b200 import synth
    fam = synth.Families(a.roles, SEED)
    kmers, roles = fam.table(int(a.table_kmers), K=a.K)
    return fam, kmers, roles


def cpu_baseline(a, fam, kmers, roles, threads=None):
    """Java-shaped oracle on the host cores over a bounded sample of the workload."""
    import oracle
    threads = threads or (os.cpu_count() or 1)
    res, off, _ = fam.batch(10_000_000, a.cpu_genomes, n_prot=N_PROT, K=a.K, mode=a.mode)
    probes = oracle.count_probes(off, a.K)
    t0 = time.time()
    db = oracle.OracleDb(kmers, roles, a.K, file_len_bytes=len(roles) * (a.K + 10), threads=threads)
    t_load = time.time() - t0
    t0 = time.time()
    out = db.apply(res, off, a.min_hits, threads=threads)
    dt = time.time() - t0
    n_seq = off.shape[0] - 1
    # the reference `apply` itself is single-threaded (ApplyKmerProcessor.java:118): same oracle, one
    # thread, on a tenth of the sample
    n1 = max(N_PROT, (n_seq // 10) // N_PROT * N_PROT)
    t0 = time.time()
    db.apply(res[: int(off[n1])], off[: n1 + 1], a.min_hits, threads=1)
    dt1 = time.time() - t0
    one = {"value": n1 / dt1, "unit": "sequences/s", "cores": 1, "sample": f"{n1} proteins, {dt1:.2f} s"}
    # "best reasonable CPU" context line: packed 64-bit keys, open addressing, all cores
    best = None
    if not a.no_cpu_best:
        t0 = time.time()
        fdb = oracle.FastDb(kmers, roles, a.K)
        t_fl = time.time() - t0
        t0 = time.time()
        fout = fdb.apply(res, off, a.min_hits, threads=threads)
        dtf = time.time() - t0
        best = {"value": n_seq / dtf, "unit": "sequences/s", "probes_per_s": probes / dtf, "cores": threads,
                "sample": f"same sample, packed-integer C port (oracle/ka_oracle_fast.c), {dtf:.2f} s; table build {t_fl:.1f} s not timed",
                "matches_java_shaped_oracle": bool(all(np.array_equal(x, y) for x, y in zip(fout, out)))}
        del fdb
    return {"value": n_seq / dt, "unit": "sequences/s", "probes_per_s": probes / dt, "cores": threads,
            "one_thread": one, "best_cpu_packed": best,
            "note": "all three lines are C proxies of the Java path, not a JVM run (no JVM in this image)",
            "kind": "port",
            "sample": f"{a.cpu_genomes} proteomes ({n_seq} proteins, {probes} probes) of the same generator, "
                      f"Java-shaped C oracle (String keys, HashMap/HashSet restatement), {threads} threads, "
                      f"{dt:.2f} s; DB load {t_load:.1f} s not timed",
            "seconds": dt}, out


def run_reference(a):
    """--impl reference: no JVM and un-vendored Maven deps => the reference cannot run here;
    its CPU implementation is represented by the oracle port on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    fam, kmers, roles = make_table(a)
    import oracle
    threads = os.cpu_count() or 1
    db = oracle.OracleDb(kmers, roles, a.K, file_len_bytes=len(roles) * (a.K + 10), threads=threads)
    res, off, _ = fam.batch(10_000_000, a.cpu_genomes, n_prot=N_PROT, K=a.K, mode=a.mode)
    probes = oracle.count_probes(off, a.K)
    n_seq = off.shape[0] - 1
    for _ in range(a.warmup):
        db.apply(res, off, a.min_hits, threads=threads)
    t0 = time.time()
    for _ in range(a.steps):
        db.apply(res, off, a.min_hits, threads=threads)
    dt = (time.time() - t0) / a.steps
    val = n_seq / dt
    line = {
        "impl": "reference", "metric": "sequences/sec", "value": val, "unit": "sequences/s",
        "probes_per_s": probes / dt, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": {"workload": workload_name(a), "sample_per_step": f"{a.cpu_genomes} proteomes"},
        "cpu_baseline": {"value": val, "unit": "sequences/s", "cores": threads, "kind": "port",
                         "sample": f"{a.cpu_genomes} proteomes per step ({n_seq} proteins, {probes} probes); "
                                   "Java-shaped C oracle: the reference is Java with un-vendored "
                                   "dependencies and no JVM exists in this image"},
        "e2e": {"value": val, "unit": "sequences/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
        return

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        # NCCL's version banner goes to stdout (NCCL_DEBUG=VERSION); stdout carries only the JSON line
        if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):
            os.environ["NCCL_DEBUG"] = "WARN"
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist:
            dist.barrier()

    def max_over_ranks(x):
        if not dist:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if not dist:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    import kmers_anno_b200 as ka
    from kmers_anno_b200.engine import pinned_array

    numa = bind_to_gpu_cpus(local)   # before any pinned allocation: first touch decides the NUMA node
    t_setup = time.time()
    fam, kmers, roles = make_table(a)
    eng = ka.Engine([local])
    for o in a.option:
        k, v = o.split("=")
        eng.set_option(k, float(v))
    eng.db_load(kmers, roles, a.K)
    info = eng.db_info()
    # this rank's shard: genomes [rank*G, (rank+1)*G), generated straight into pinned memory
    res, off, _ = fam.batch(rank * a.genomes, a.genomes, n_prot=N_PROT, K=a.K, mode=a.mode, alloc=pinned_array)
    n_seq = off.shape[0] - 1
    out = (pinned_array(n_seq, np.int32), pinned_array(n_seq, np.int32), pinned_array(n_seq, np.uint8))
    t_setup = time.time() - t_setup

    # ---- resident loop: value --------------------------------------------------------
    batch = eng.upload(res, off)
    probes = eng.stats()["probes"]
    sampler = ClockSampler(local)
    sampler.start()          # nvidia-smi needs ~1 s to produce its first line: start before the warm-up
    for _ in range(a.warmup):
        eng.annotate_resident(batch, a.min_hits)
    barrier()
    sampler.mark()           # samples from here on are inside the timed regions
    k_ms, t_ms, launches = 0.0, 0.0, 0
    for _ in range(a.steps):
        eng.annotate_resident(batch, a.min_hits)   # synchronises its stream before returning
        st = eng.stats()
        k_ms += st["kernel_ms"]; t_ms += st["tile_kernel_ms"]; launches += st["kernel_launches"]
    barrier()
    dev_role, dev_hits, dev_flag = eng.download(batch)
    batch.free()
    step_ms = max_over_ranks(k_ms / a.steps)
    tile_ms = max_over_ranks(t_ms / a.steps)
    total_seq = sum_over_ranks(float(n_seq))
    total_probes = sum_over_ranks(float(probes))

    # ---- e2e loop: ka_annotate with pinned host buffers ------------------------------
    e2e = None
    if not a.no_e2e:
        for _ in range(a.warmup):
            eng.annotate(res, off, a.min_hits, out=out)
        barrier()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            eng.annotate(res, off, a.min_hits, out=out)   # returns after the D2H of the results
        e2e_ms = (time.perf_counter() - t0) * 1e3 / a.steps
        barrier()
        st = eng.stats()
        e2e_ms = max_over_ranks(e2e_ms)
        same = bool(np.array_equal(out[0], dev_role) and np.array_equal(out[1], dev_hits)
                    and np.array_equal(out[2], dev_flag))
        e2e = {"value": total_seq / (e2e_ms * 1e-3), "unit": "sequences/s",
               "probes_per_s": total_probes / (e2e_ms * 1e-3), "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": int(st["h2d_bytes"]), "d2h_bytes_per_step": int(st["d2h_bytes"]),
               "host_memory": "pinned (ka_host_alloc)", "matches_resident_results": same}

    clocks = sampler.stop()

    # ---- config 2 as stated: ONE proteome per call (latency-bound, reported for completeness) ----
    c2 = None
    if rank == 0:
        n1 = N_PROT
        r1 = res[: int(off[n1])]
        o1 = off[: n1 + 1]
        out1 = tuple(x[:n1] for x in out)
        b1 = eng.upload(r1, o1)
        p1 = eng.stats()["probes"]
        for _ in range(5):
            eng.annotate_resident(b1, a.min_hits)
        ks = []
        for _ in range(50):
            eng.annotate_resident(b1, a.min_hits)
            ks.append(eng.stats()["kernel_ms"])
        b1.free()
        for _ in range(5):
            eng.annotate(r1, o1, a.min_hits, out=out1)
        t0 = time.perf_counter()
        for _ in range(50):
            eng.annotate(r1, o1, a.min_hits, out=out1)
        e1 = (time.perf_counter() - t0) / 50 * 1e3
        k1 = float(np.median(ks))
        c2 = {"workload": "C2: one 4,500-protein proteome per call against the same table",
              "kernel_ms": k1, "kernel_probes_per_s": p1 / (k1 * 1e-3), "e2e_ms": e1,
              "e2e_sequences_per_s": n1 / (e1 * 1e-3), "probes": int(p1)}

    # ---- roofline of the dominant kernel ----------------------------------------------
    peak, peak_src = measured_peak()
    achieved = BYTES_PER_PROBE * probes / (tile_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "tile_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": NCU_DRAM_BYTES_PER_PROBE * probes,
                "traffic_note": "ncu dram bytes per probe (95.6 B, 60-proteome capture) x probes of this launch",
                "peak_source": peak_src,
                "algorithmic_bytes_per_probe": BYTES_PER_PROBE, "probes_per_launch": int(probes),
                "kernel_ms": tile_ms}
    rand = None
    if rank == 0:
        r_rand = eng.probe_roofline(info["table_bytes"], 1 << 28, slot_bytes=32, reps=5)
        pps = probes / (tile_ms * 1e-3)
        rand = {"r_rand_probes_per_s": r_rand, "r_rand_GBps_at_32B": r_rand * 32 / 1e9,
                "buffer_bytes": int(info["table_bytes"]), "achieved_probes_per_s": pps, "frac": pps / r_rand}

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cpu, cpu_out = cpu_baseline(a, fam, kmers, roles)
        # the same sample on the GPU must give the oracle's answer (checker, not the product)
        s_res, s_off, _ = fam.batch(10_000_000, a.cpu_genomes, n_prot=N_PROT, K=a.K, mode=a.mode)
        g = eng.annotate(s_res, s_off, a.min_hits)
        cpu["gpu_matches_oracle_on_sample"] = bool(all(np.array_equal(x, y) for x, y in zip(g, cpu_out)))
    eng.close()

    if rank == 0:
        line = {
            "metric": "sequences/sec", "value": total_seq / (step_ms * 1e-3), "unit": "sequences/s",
            "probes_per_s": total_probes / (step_ms * 1e-3),
            "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": step_ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic",
            "config": {"workload": workload_name(a), "sequences_per_gpu": int(n_seq),
                       "residues_per_gpu": int(off[-1]), "probes_per_gpu": int(probes), "K": a.K,
                       "min_hits": a.min_hits, "table_bytes": int(info["table_bytes"]),
                       "table_keys": int(info["n_keys"]), "parallelism": f"replicated table, {world} shard(s)",
                       "l2": "table (>> 126 MB L2) probed at random and batch residues > L2: no flush needed",
                       "setup_s": round(t_setup, 1), "host_affinity": numa},
            "e2e": e2e, "gpu_launches": int(launches) * world, "clocks": clocks,
            "roofline": roofline, "rand_roofline": rand, "cpu_baseline": cpu, "c2_single_proteome": c2,
        }
        print(json.dumps(line), flush=True)
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
