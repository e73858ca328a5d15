#!/usr/bin/env python
"""bench.py -- demuxlet droplet-likelihood throughput (BASELINE.json metric) on N B200s.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's CPU algorithm (oracle port) on host cores

A "step" is one pass of the hot path over one batch of synthetic droplets: every barcode of the
workload scored against all samples and all ordered sample pairs x alphas.  Workload at any N:
BASELINE configs[1] ("demuxlet 1xB200: 64 samples, 200k SNPs, 20k barcodes, alpha grid 0-0.5 step
0.05", ~200 SNP-reads per cell -- our assumption, printed in config) PER GPU: barcodes are sharded
across ranks with the genotype matrix replicated and no data-path collective (weak scaling).

  value  : droplets/s with the CSR store, genotype matrix and tiles resident in HBM
           (K1 tile + compaction + K2s singlet + K2 score + K3 finalize), CUDA events, max over ranks.
  e2e    : droplets/s through the host-buffer C-ABI calls a patched cmdCramDemuxlet makes per run
           (ccu_demux_set_genotypes + ccu_demux_score: pinned host -> device copies, all kernels,
           result copy back), wall clock around synchronised calls, max over ranks.
  roofline: the scoring kernel K2 against the FP64 pipe (the path is FP64-FMA bound).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "droplets/sec (all sample-pair x alpha LLKs)"
UNIT = "droplets/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C2", help="BASELINE config shape (C1..C5); default C2")
    ap.add_argument("--nbcs", type=int, default=None, help="override barcodes per GPU (debugging)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="budget of the cpu_baseline leg")
    return ap.parse_args()


def config_dict(d, n_gpus):
    cfg = d["cfg"]
    return {"workload": "%s demuxlet: %d samples, %d SNPs, %d barcodes per GPU, %d alphas, ~%d SNP-reads/cell "
                        "(r-bar is our assumption), %.0f%% doublets" %
                        (d["name"], cfg["nv"], cfg["nsnps"], cfg["nbcs"], len(cfg["alphas"]), cfg["rbar"],
                         100 * cfg["doublet_rate"]),
            "samples": cfg["nv"], "snps": cfg["nsnps"], "barcodes_per_gpu": cfg["nbcs"],
            "alphas": len(cfg["alphas"]), "reads_per_cell": cfg["rbar"], "entries_per_gpu": int(d["cell_ptr"][-1]),
            "seed": d["seed"], "parallelism": "barcode-shard x%d, genotypes replicated" % n_gpus,
            "l2_policy": "inputs larger than L2 (FP64 genotype matrix + tiles >> 126 MB); no explicit flush"}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe):
    one long-running `nvidia-smi -lms 100` whose rows are cut to the timed window."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.t0, self.t1 = [], None, None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.strip().split(",")]
            if len(f) >= 7:
                self.rows.append((time.perf_counter(), f))

    def begin(self):
        self.t0 = time.perf_counter()

    def end(self):
        self.t1 = time.perf_counter()

    def summary(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [f for (t, f) in self.rows if self.t0 is not None and self.t0 <= t <= (self.t1 or 1e30) + 0.1]
        if not rows:
            rows = [f for (_, f) in self.rows[-3:]]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no nvidia-smi sample"]}
        sm = sorted(float(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": float(rows[0][1]), "samples": len(sm),
                "power_w_max": max(float(r[2]) for r in rows), "reasons": reasons}


REF_BIN = os.path.join(ROOT, "oracle", "_ref", "ref_cramore")


def subset_dataset(d, begin, end):
    """Barcodes [begin, end) of a synthetic dataset with the SNP panel reduced to the SNPs they touch
    (the per-droplet cost of the reference does not depend on the panel size, its load time does)."""
    cp, rp = np.asarray(d["cell_ptr"], np.int64), np.asarray(d["read_ptr"], np.int64)
    e0, e1 = int(cp[begin]), int(cp[end])
    r0, r1 = int(rp[e0]), int(rp[e1])
    snps = np.unique(np.asarray(d["snp_idx"][e0:e1]))
    remap = np.full(d["nsnps"], -1, np.int64)
    remap[snps] = np.arange(len(snps))
    return dict(cell_ptr=cp[begin:end + 1] - e0, snp_idx=remap[np.asarray(d["snp_idx"][e0:e1])].astype(np.int32),
                read_ptr=rp[e0:e1 + 1] - r0, al=np.asarray(d["al"][r0:r1]), bq=np.asarray(d["bq"][r0:r1]),
                af=np.asarray(d["af"])[snps], geno=np.asarray(d["geno"])[snps], nsnps=len(snps), nv=d["nv"],
                sample_ids=d["sample_ids"], alphas=d["alphas"], barcodes=None)


def reference_rate(d, seconds, procs):
    """The reference's OWN demuxlet engine (oracle/_ref/ref_cramore: /root/reference/cmd_cram_demuxlet.cpp and
    sc_drop_seq.cpp compiled unmodified behind the zlib I/O shim) on `procs` disjoint barcode subsets run as
    parallel processes -- the reference is single-threaded and scales by --group-list style sharding
    (cmd_cram_demuxlet.cpp:75).  Returns (droplets/s, droplets, wall seconds)."""
    import shutil
    import tempfile
    from cramore_b200.synth import write_plp, write_vcf
    nv, nA = d["nv"], len(d["alphas"])
    nnz_cell = float(d["cell_ptr"][-1]) / max(1, d["nbcs"])
    est = nnz_cell * nv * nv * nA * 9 * 1.3e-9 + 1e-4  # ~seconds per droplet on one core (9 FMA + log share)
    per = int(max(2, min(d["nbcs"] // procs, seconds / est)))
    tmp = tempfile.mkdtemp(prefix="ccu_ref_")
    try:
        cmds = []
        for p in range(procs):
            sub = subset_dataset(d, p * per, (p + 1) * per)
            pre = os.path.join(tmp, "p%d" % p)
            write_plp(sub, pre)
            write_vcf(sub, pre + ".vcf")
            cmd = [REF_BIN, "demuxlet", "--plp", pre, "--vcf", pre + ".vcf", "--field", "GP", "--out", pre + ".out",
                   "--min-mac", "0", "--min-callrate", "0"]
            for a in d["alphas"]:
                cmd += ["--alpha", repr(float(a))]
            cmds.append(cmd)
        t0 = time.perf_counter()
        ps = [subprocess.Popen(c, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for c in cmds]
        rcs = [p.wait() for p in ps]
        dt = time.perf_counter() - t0
        if any(rcs):
            raise RuntimeError("ref_cramore failed: %s" % rcs)
        rows = sum(sum(1 for _ in open(os.path.join(tmp, "p%d.out.best" % p))) - 1 for p in range(procs))
        assert rows == per * procs, (rows, per, procs)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return per * procs / dt, per * procs, dt


def cpu_oracle_rate(d, seconds, threads):
    """Reference algorithm (oracle port of cmd_cram_demuxlet.cpp:636-906) on the host: droplets/s over
    a bounded barcode prefix, `threads` workers over disjoint barcode ranges (--group-list recipe)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import pyoracle
    gps = pyoracle.mix_genotypes(d["gp"], d["snp_err"])
    args = (d["cell_ptr"], d["snp_idx"], d["read_ptr"], d["al"], d["bq"], gps, d["alphas"], 0.5)
    t0 = time.perf_counter()
    pyoracle.demux_score(*args, cell_begin=0, cell_end=2)
    per_cell = max((time.perf_counter() - t0) / 2, 1e-6)
    per_thread = int(max(2, min(d["nbcs"] // threads, seconds / per_cell)))
    ranges = [(i * per_thread, (i + 1) * per_thread) for i in range(threads)]
    t0 = time.perf_counter()
    if threads == 1:
        pyoracle.demux_score(*args, cell_begin=ranges[0][0], cell_end=ranges[0][1])
    else:
        with ThreadPoolExecutor(threads) as ex:  # ctypes releases the GIL during the foreign call
            list(ex.map(lambda r: pyoracle.demux_score(*args, cell_begin=r[0], cell_end=r[1]), ranges))
    dt = time.perf_counter() - t0
    n = per_thread * threads
    return n / dt, n, dt


def cpu_rate(d, seconds, threads):
    """(rate, n, seconds, kind): the reference's own compiled engine when oracle/_ref travelled here, else the
    oracle port."""
    if os.path.exists(REF_BIN):
        r, n, dt = reference_rate(d, seconds, threads)
        return r, n, dt, "reference"
    r, n, dt = cpu_oracle_rate(d, seconds, threads)
    return r, n, dt, "port"


def run_reference(args):
    """--impl reference: the reference's own CPU implementation timed on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from cramore_b200.synth import make_dataset
    d = make_dataset(args.workload, nbcs=args.nbcs)
    threads = os.cpu_count() or 1
    per_step = max(2.0, min(10.0, 150.0 / max(1, args.steps + args.warmup)))
    rates, n_last, kind = [], 0, "port"
    for i in range(args.warmup + args.steps):
        r, n_last, dt, kind = cpu_rate(d, per_step, threads)
        if i >= args.warmup:
            rates.append((n_last, dt))
    tot_n = sum(n for n, _ in rates)
    tot_t = sum(t for _, t in rates)
    value = tot_n / tot_t
    sample = "%d barcodes per step (%d per process, %d processes) of the same %s workload, SNP panel reduced to " \
             "the SNPs those barcodes touch" % (n_last, n_last // threads, threads, d["name"])
    note = ("oracle/_ref/ref_cramore: the reference's cmd_cram_demuxlet.cpp + sc_drop_seq.cpp compiled unmodified "
            "behind a zlib I/O shim (htslib absent), one process per core over disjoint barcode sets"
            if kind == "reference" else "oracle port of cmd_cram_demuxlet.cpp:636-906 (oracle/_ref not present)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(1, args.steps),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(d, args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "note": note}
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa_node(local):
    """Run this process (and therefore place the pinned host buffers it allocates next) on the CPUs of the NUMA node
    the GPU hangs off: host->device copies from the far socket run at about half the PCIe rate.  Returns a short
    description for the JSON line; a no-op where sysfs / NVML do not say (or BENCH_NO_NUMA_BIND=1)."""
    if os.environ.get("BENCH_NO_NUMA_BIND") == "1":
        return "off (BENCH_NO_NUMA_BIND=1)"
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = local
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if local < len(ids) and ids[local].isdigit():
                idx = int(ids[local])
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(idx)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.split(":", 1)
        path = "/sys/bus/pci/devices/%s:%s" % (dom[-4:].lower(), rest.lower())
        node = int(open(path + "/numa_node").read().strip())
        if node < 0:
            return "single node (numa_node=-1)"
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return "node %d has no allowed CPU" % node
        os.sched_setaffinity(0, cpus)
        return "numa node %d (%d cpus)" % (node, len(cpus))
    except Exception as e:  # noqa: BLE001 -- placement is an optimisation, never a reason to fail the bench
        return "unavailable (%s)" % type(e).__name__


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from cramore_b200 import RESULT_DTYPE, DemuxEngine
    from cramore_b200.synth import algorithmic_flops_per_entry, make_dataset

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the engine has no CPU fallback")
    torch.cuda.set_device(local)
    host_affinity = bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- synthetic workload: same genotypes on every rank, rank-specific droplets (weak scaling)
    d = make_dataset(args.workload, nbcs=args.nbcs, droplet_seed_offset=rank)
    nbcs, nnz = d["nbcs"], int(d["cell_ptr"][-1])

    def pinned(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t, t.numpy()

    keep = []
    host = {}
    for k in ("gp", "snp_err", "cell_ptr", "snp_idx", "read_ptr", "al", "bq"):
        t, host[k] = pinned(d[k])
        keep.append(t)
    out_t = torch.empty(nbcs * RESULT_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
    out = out_t.numpy().view(RESULT_DTYPE)
    h2d = sum(host[k].nbytes for k in host)
    d2h = out.nbytes

    eng = DemuxEngine(local)
    # one real stream shared by the engine and the torch events (handle 0, the legacy default stream,
    # would mean "engine-owned stream" to ccu_set_stream)
    stream = torch.cuda.Stream(device=torch.device("cuda", local))
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)
    eng.configure(d["nv"], d["alphas"], 0.5)

    e2e_parts = [0.0, 0.0]  # seconds inside set_genotypes / score (host wall clock)
    e2e_dev = [0.0, 0.0, 0.0, 0.0]  # the engine's own CUDA-event times of the same calls (ms)

    def e2e_step():
        t_a = time.perf_counter()
        eng.set_genotypes(host["gp"], host["snp_err"])
        t_b = time.perf_counter()
        eng.score(host["cell_ptr"], host["snp_idx"], host["read_ptr"], host["al"], host["bq"], out)
        e2e_parts[0] += t_b - t_a
        e2e_parts[1] += time.perf_counter() - t_b
        tm = eng.timings()
        for i, k in enumerate(("ms_genotypes", "ms_h2d", "ms_run", "ms_d2h")):
            e2e_dev[i] += tm[k]

    for _ in range(max(args.warmup, 3)):
        e2e_step()
    fp64_peak = eng.fp64_peak_tflops(300.0)

    # ---- leg 1: HBM-resident inputs, device time
    eng.set_genotypes(host["gp"], host["snp_err"])
    eng.upload(host["cell_ptr"], host["snp_idx"], host["read_ptr"], host["al"], host["bq"])
    eng.run()
    eng.wait()
    sampler = ClockSampler(local)
    time.sleep(0.3)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k2_ms, k1_ms, launches = [], [], 0
    barrier()
    sampler.begin()
    ev0.record(stream)
    for _ in range(args.steps):
        eng.run()
        eng.wait()  # per-step event readout; the stream stays busy, the host sync is ~10 us per step
        t = eng.timings()
        k2_ms.append(t["ms_score"])
        k1_ms.append(t["ms_tile"])
        launches += t["launches"]
    ev1.record(stream)
    barrier()
    sampler.end()
    dev_ms = ev0.elapsed_time(ev1)
    clocks = sampler.summary()
    dev_ms_max = max_over_ranks(dev_ms)
    total_droplets = sum_over_ranks(nbcs) * args.steps
    value = total_droplets / (dev_ms_max * 1e-3)

    # ---- leg 2: end to end through the host-buffer C ABI
    barrier()
    e2e_parts[0] = e2e_parts[1] = 0.0
    e2e_dev[:] = [0.0, 0.0, 0.0, 0.0]
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
        launches += eng.timings()["launches"] + 2
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = total_droplets / e2e_s
    res_check = np.array(out[:4])

    # ---- roofline of the dominant kernel (K2 scoring, FP64 pipe)
    nv, nA = d["nv"], len(d["alphas"])
    flops_launch = algorithmic_flops_per_entry(nv, list(d["alphas"])) * nnz
    k2_avg = float(np.mean(k2_ms))
    achieved = flops_launch / (k2_avg * 1e-3) / 1e12
    # FP64 instructions the kernel really issues: every ordered pair of the padded sample
    # tiles.  Affine entries (<= 1 informative read) go two SNPs at a time: a thread tile of 2 j x TK k x AC
    # alphas costs 2 + TK*(10 + 6*AC) instructions per SNP pair (1.8 per likelihood and SNP at AC=10, TK=1;
    # the odd SNP of a run costs 2.1, not counted); general entries 4 per (j,k,alpha) + 9 per (k,alpha) per
    # j-tile for T
    al_ne2 = (np.asarray(d["al"]) != 2).astype(np.int64)
    csum = np.concatenate([[0], np.cumsum(al_ne2)])
    neff = csum[np.asarray(d["read_ptr"][1:])] - csum[np.asarray(d["read_ptr"][:-1])]
    n_gen = int((neff > 1).sum())
    n_aff = nnz - n_gen
    nvp = (nv + 31) // 32 * 32
    nd = nA - 1
    ac = min((10, 5, 2, 1), key=lambda a: ((nd + a - 1) // a * a, -a))
    nd_pad = (nd + ac - 1) // ac * ac
    nac = nd_pad // ac
    tk = {10: 1, 5: 2, 2: 4, 1: 8}[ac]
    instr_aff = (nvp // 2) * (nvp // tk) * nac * (1 + tk * (5 + 3 * ac))
    instr_gen = nvp * nvp * 4 * nd_pad + (nvp // 32) * nvp * nd_pad * 9
    executed = 2.0 * (n_aff * instr_aff + n_gen * instr_gen)
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01_k2_traffic.json")))
        if d["name"] == "C2" and nbcs == 20000:
            traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
    except Exception:
        pass
    roofline = {"bound": "fp64", "kernel": "demux_score_kernel (K2)", "achieved": achieved, "peak": fp64_peak,
                "unit": "TFLOP/s", "frac": achieved / fp64_peak if fp64_peak > 0 else None, "traffic": traffic,
                "traffic_note": "dram__bytes_read+write of one K2 launch at this workload, ncu --set full capture "
                                "summarised in profiles/r01_k2_score_ncu_full.csv (bytes; the kernel is FP64-pipe "
                                "bound, HBM is at ~1% of peak)",
                "peak_source": "DFMA microkernel measured in this run (MEASURED_PEAKS.json has no FP64 entry; "
                               "nominal 148 SM x 64 DFMA/clk x 2 x 1.965 GHz = 37.2)",
                "note": "achieved = SURVEY 8(d) algorithmic flops (8 per needed LLK + 18*nv*nA per entry) / K2 event "
                        "time. The affine fast path needs ~half of those instructions, so frac can exceed 1; "
                        "executed_frac counts the FP64 instructions actually issued (x2 flops) and is the pipe "
                        "utilisation ncu reports as sm__pipe_fp64_cycles_active",
                "algorithmic_flops_per_launch": flops_launch, "executed_flops_per_launch": executed,
                "executed_frac": executed / (k2_avg * 1e-3) / 1e12 / fp64_peak if fp64_peak > 0 else None,
                "entries_affine": n_aff, "entries_general": n_gen,
                "k2_ms_avg": k2_avg, "k2_share_of_step": k2_avg * args.steps / dev_ms}
    tile_bytes = 2 * int(d["read_ptr"][-1]) + 8 * nnz + 33 * nnz + 80 * len(d["alphas"]) * n_gen  # reads + read_ptr + vaff/eclass + general tiles
    k1_avg = float(np.mean(k1_ms))
    hbm_peak = None
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        hbm_peak = 6650.0
    roofline_k1 = {"bound": "hbm", "kernel": "demux_entry_kernel + demux_tile_kernel (K1)", "achieved": tile_bytes / (k1_avg * 1e-3) / 1e9,
                   "peak": hbm_peak, "unit": "GB/s", "frac": tile_bytes / (k1_avg * 1e-3) / 1e9 / hbm_peak,
                   "k1_ms_avg": k1_avg}

    line = None
    if rank == 0:
        cpu_val, cpu_n, cpu_dt, cpu_kind = cpu_rate(d, args.cpu_seconds, 1)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config_dict(d, world), "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": 1e3 * e2e_s / args.steps,
                        "ms_set_genotypes_call": 1e3 * e2e_parts[0] / args.steps,
                        "ms_score_call": 1e3 * e2e_parts[1] / args.steps, "host_affinity": host_affinity,
                        "device_ms_inside": {k: v / args.steps for k, v in zip(
                            ("genotype_mix_kernels", "csr_h2d", "kernels", "result_d2h"), e2e_dev)},
                        "includes": "genotype matrix upload + mixing, CSR upload, all kernels, result download"},
                "gpu_launches": launches, "roofline": roofline, "roofline_k1": roofline_k1,
                "cpu_baseline": {"value": cpu_val, "unit": UNIT, "cores": 1, "kind": cpu_kind,
                                 "sample": "first %d barcodes of the same workload, %.1f s, single thread "
                                           "(the reference is single-threaded); kind=reference is "
                                           "oracle/_ref/ref_cramore, the reference's own engine sources compiled "
                                           "behind a zlib I/O shim" % (cpu_n, cpu_dt)},
                "kernel_ms": eng.timings(), "sample_result": [float(res_check["sngBestLLK"][0]),
                                                              float(res_check["dblBestLLK"][0])]}
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
