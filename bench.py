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
  extra  : the other named BASELINE configs measured in the same run, so that they carry the driver's clock record:
           extra.c3 (256 samples x 500k SNPs, 12.5k barcodes per GPU = the 100k-barcode 8-GPU target at N=8),
           extra.c5 (high depth), extra.c2_skew (Zipf SNP draw: the >= 2-reads-per-SNP path of K2) and
           extra.fmx_c4 (freemuxlet EM, K=16, 50k barcodes, 300k SNPs, sharded over the N ranks with the
           peer-memory push exchange over NVLink and with the NCCL all-reduce exchange).  --no-extras skips them.
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
    ap.add_argument("--no-extras", action="store_true", help="only the headline workload (no extra.* blocks)")
    ap.add_argument("--extras", default="c3,c5,c2_skew,fmx_c4", help="comma-separated extra blocks to run")
    ap.add_argument("--extras-deadline", type=float, default=420.0,
                    help="seconds after which the headline line is printed without the unfinished extras")
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



class Ctx:
    """Process-wide handles of one bench run: torch, the process group, the one CUDA stream shared by the engine,
    torch and NCCL, and the cross-rank reductions."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a B200: the engine has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.host_affinity = bind_to_gpu_numa_node(self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        # one real stream shared by the engine and the torch events (handle 0, the legacy default stream,
        # would mean "engine-owned stream" to ccu_set_stream)
        self.stream = torch.cuda.Stream(device=self.dev)
        torch.cuda.set_stream(self.stream)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def _reduce(self, x, op):
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max_over_ranks(self, x):
        return self._reduce(x, self.dist.ReduceOp.MAX)

    def min_over_ranks(self, x):
        return self._reduce(x, self.dist.ReduceOp.MIN)

    def sum_over_ranks(self, x):
        return self._reduce(x, self.dist.ReduceOp.SUM)

    def pinned(self, a):
        t = self.torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t, t.numpy()


class _DevArray:
    """CUDA array interface view of engine-owned device memory (no copy)."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def k2_instruction_model(d):
    """FP64 instructions the scoring kernel really issues per launch, x2 = executed flops.  Every
    ordered pair of the padded sample tiles.  Affine entries (<= 1 informative read) go two SNPs at a time: a thread
    tile of 2 j x TK k x AC alphas costs 2 + TK*(10 + 6*AC) instructions per SNP pair (1.8 per likelihood and SNP at
    AC=10, TK=1; the odd SNP of a run costs 2.1, not counted); general entries 4 per (j,k,alpha) + 9 per (k,alpha)
    per j-tile for T."""
    nv, nA = d["nv"], len(d["alphas"])
    nnz = int(d["cell_ptr"][-1])
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
    return 2.0 * (n_aff * instr_aff + n_gen * instr_gen), n_aff, n_gen


def measure_demux(ctx, d, steps, warmup, fp64_peak=None):
    """Both legs of the demuxlet measurement on the workload `d` (this rank's shard; weak scaling):
    device-resident (CUDA events, max over ranks) and end to end from pinned host buffers through the C ABI."""
    torch, dist = ctx.torch, ctx.dist
    from cramore_b200 import RESULT_DTYPE, DemuxEngine
    from cramore_b200.synth import algorithmic_flops_per_entry
    world, rank, local, stream = ctx.world, ctx.rank, ctx.local, ctx.stream
    nbcs, nnz = d["nbcs"], int(d["cell_ptr"][-1])
    nsnps, nv = d["nsnps"], d["nv"]
    keep, host = [], {}
    for k in ("gp", "snp_err", "cell_ptr", "snp_idx", "read_ptr", "al", "bq"):
        t, host[k] = ctx.pinned(d[k])
        keep.append(t)
    host_t = dict(zip(("gp", "snp_err", "cell_ptr", "snp_idx", "read_ptr", "al", "bq"), keep))
    out_t = torch.empty(nbcs * RESULT_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
    out = out_t.numpy().view(RESULT_DTYPE)
    csr_bytes = sum(host[k].nbytes for k in ("cell_ptr", "snp_idx", "read_ptr", "al", "bq"))

    eng = DemuxEngine(local)
    eng.set_stream(stream.cuda_stream)
    eng.configure(nv, d["alphas"], 0.5)

    # ---- genotype matrix at N > 1: ONE copy crosses PCIe per node.  Every rank copies its SNP slab of the float
    # matrix host->device into the engine's staging array and an in-place NCCL all-gather over NVLink completes it
    # (ccu_demux_genotypes_stage / _commit); at N = 1 the plain host-pointer call.
    row = nv * 3
    chunk = nsnps // world
    gathered = None
    if world > 1:
        gathered = torch.empty(world * nbcs * RESULT_DTYPE.itemsize, dtype=torch.uint8, device=ctx.dev)
        gathered_host = torch.empty(world * nbcs * RESULT_DTYPE.itemsize, dtype=torch.uint8).pin_memory() if rank == 0 else None
    gp_flat = host_t["gp"].view(-1)

    def set_genotypes_e2e():
        if world == 1:
            eng.set_genotypes(host["gp"], host["snp_err"])
            return host["gp"].nbytes + host["snp_err"].nbytes
        gpd, errd, _ = eng.genotypes_stage(nsnps)
        full = torch.as_tensor(_DevArray(gpd, nsnps * row, "<f4"), device=ctx.dev)
        errv = torch.as_tensor(_DevArray(errd, nsnps, "<f8"), device=ctx.dev)
        lo, hi = rank * chunk * row, (rank + 1) * chunk * row
        full[lo:hi].copy_(gp_flat[lo:hi], non_blocking=True)
        nb = (hi - lo) * 4
        if chunk > 0:
            dist.all_gather_into_tensor(full[:world * chunk * row], full[lo:hi])
        if world * chunk < nsnps:  # the few SNPs left over by the equal split: every rank copies them itself
            full[world * chunk * row:].copy_(gp_flat[world * chunk * row:], non_blocking=True)
            nb += (nsnps - world * chunk) * row * 4
        errv.copy_(host_t["snp_err"], non_blocking=True)
        eng.genotypes_commit(nsnps, False)
        return nb + host["snp_err"].nbytes

    def gather_e2e():
        """The final result gather (north_star): the shards' records all-gathered over NVLink from the engine's
        device array, the full table copied to the host on rank 0."""
        if world == 1:
            return 0
        p, n = eng.results_dev()
        mine = torch.as_tensor(_DevArray(p, n * RESULT_DTYPE.itemsize, "|u1"), device=ctx.dev)
        dist.all_gather_into_tensor(gathered, mine)
        if rank == 0:
            gathered_host.copy_(gathered, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return gathered_host.numel()
        return 0

    e2e_parts = [0.0, 0.0, 0.0]  # seconds inside set_genotypes / score / gather (host wall clock)
    e2e_dev = [0.0, 0.0, 0.0, 0.0]  # the engine's own CUDA-event times of the same calls (ms)
    e2e_bytes = [0, 0]

    def e2e_step():
        t_a = time.perf_counter()
        hb = set_genotypes_e2e()
        t_b = time.perf_counter()
        eng.score(host["cell_ptr"], host["snp_idx"], host["read_ptr"], host["al"], host["bq"], out)
        t_c = time.perf_counter()
        db = gather_e2e()
        e2e_parts[0] += t_b - t_a
        e2e_parts[1] += t_c - t_b
        e2e_parts[2] += time.perf_counter() - t_c
        e2e_bytes[0], e2e_bytes[1] = hb + csr_bytes, out.nbytes + db
        tm = eng.timings()
        for i, k in enumerate(("ms_genotypes", "ms_h2d", "ms_run", "ms_d2h")):
            e2e_dev[i] += tm[k]

    for _ in range(max(warmup, 3)):
        e2e_step()
    if fp64_peak is None:
        fp64_peak = eng.fp64_peak_tflops(300.0)

    # ---- leg 1: HBM-resident inputs, device time
    eng.upload(host["cell_ptr"], host["snp_idx"], host["read_ptr"], host["al"], host["bq"])
    eng.run()
    eng.wait()
    sampler = ClockSampler(local)
    time.sleep(0.3)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k2_ms, k1_ms, launches = [], [], 0
    ctx.barrier()
    sampler.begin()
    ev0.record(stream)
    for _ in range(steps):
        eng.run()
        eng.wait()  # per-step event readout; the stream stays busy, the host sync is ~10 us per step
        t = eng.timings()
        k2_ms.append(t["ms_score"])
        k1_ms.append(t["ms_tile"])
        launches += t["launches"]
    ev1.record(stream)
    ctx.barrier()
    dev_ms = ev0.elapsed_time(ev1)
    dev_ms_max = ctx.max_over_ranks(dev_ms)
    total_droplets = ctx.sum_over_ranks(nbcs) * steps
    value = total_droplets / (dev_ms_max * 1e-3)
    kernel_ms = eng.timings()

    # ---- leg 2: end to end through the host-buffer C ABI (+ the final result gather at N > 1)
    ctx.barrier()
    e2e_parts[:] = [0.0, 0.0, 0.0]
    e2e_dev[:] = [0.0, 0.0, 0.0, 0.0]
    t0 = time.perf_counter()
    for _ in range(steps):
        e2e_step()
        launches += eng.timings()["launches"] + 2
    ctx.barrier()
    e2e_s = ctx.max_over_ranks(time.perf_counter() - t0)
    sampler.end()
    clocks = sampler.summary()
    e2e_value = total_droplets / e2e_s
    res_check = np.array(out[:4])

    # ---- roofline of the dominant kernel (K2 scoring, FP64 pipe)
    flops_survey = algorithmic_flops_per_entry(nv, list(d["alphas"])) * nnz
    executed, n_aff, n_gen = k2_instruction_model(d)
    k2_avg = float(np.mean(k2_ms))
    achieved = executed / (k2_avg * 1e-3) / 1e12
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "k2_traffic.json")))
        if d["name"] == tj.get("workload", "C2") and nbcs == tj.get("nbcs", 20000) and not d.get("snp_skew"):
            traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
    except Exception:
        pass
    roofline = {"bound": "fp64", "kernel": "demux_score_kernel (K2)", "achieved": achieved, "peak": fp64_peak,
                "unit": "TFLOP/s", "frac": achieved / fp64_peak if fp64_peak > 0 else None, "traffic": traffic,
                "traffic_note": "dram__bytes_read+write of one K2 launch at this workload, ncu --set full capture "
                                "summarised under profiles/ (bytes; the kernel is FP64-pipe bound, HBM is at ~1% of peak)",
                "peak_source": "DFMA microkernel measured in this run (MEASURED_PEAKS.json has no FP64 entry; "
                               "nominal 148 SM x 64 DFMA/clk x 2 x 1.965 GHz = 37.2)",
                "note": "achieved = FP64 flops of the algorithm as built (DESIGN.md 3: 1.8 FP64 instructions per "
                        "likelihood and SNP on single-read entries through the exact affine identity, 4 + T on "
                        "entries with >= 2 reads, every ordered pair of the padded sample tiles) / K2 event time; "
                        "frac is therefore the FP64-pipe utilisation ncu reports as sm__pipe_fp64_cycles_active. "
                        "SURVEY 8(d) charges 8 flops per LLK: that count / time is survey_algorithmic_tflops",
                "algorithmic_flops_per_launch": executed, "survey_flops_per_launch": flops_survey,
                "survey_algorithmic_tflops": flops_survey / (k2_avg * 1e-3) / 1e12,
                "algorithmic_speedup_vs_survey_count": flops_survey / executed if executed > 0 else None,
                "entries_affine": n_aff, "entries_general": n_gen,
                "k2_ms_avg": k2_avg, "k2_share_of_step": k2_avg * steps / dev_ms}
    tile_bytes = 2 * int(d["read_ptr"][-1]) + 8 * nnz + 33 * nnz + 80 * len(d["alphas"]) * n_gen  # reads + read_ptr + vaff/eclass + general tiles
    k1_avg = float(np.mean(k1_ms))
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        hbm_peak = 6650.0
    roofline_k1 = {"bound": "hbm", "kernel": "demux_entry_kernel + demux_tile_kernel (K1)", "achieved": tile_bytes / (k1_avg * 1e-3) / 1e9,
                   "peak": hbm_peak, "unit": "GB/s", "frac": tile_bytes / (k1_avg * 1e-3) / 1e9 / hbm_peak,
                   "k1_ms_avg": k1_avg}
    eng.close()
    del keep, host_t
    return {"value": value, "ms_per_step": dev_ms_max / steps, "clocks": clocks, "fp64_peak": fp64_peak,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(e2e_bytes[0]),
                    "d2h_bytes_per_step": int(e2e_bytes[1]),
                    "ms_per_step": 1e3 * e2e_s / steps,
                    "ms_set_genotypes_call": 1e3 * e2e_parts[0] / steps,
                    "ms_score_call": 1e3 * e2e_parts[1] / steps,
                    "ms_result_gather": 1e3 * e2e_parts[2] / steps, "host_affinity": ctx.host_affinity,
                    "device_ms_inside": {k: v / steps for k, v in zip(
                        ("genotype_mix_kernels", "csr_h2d", "kernels", "result_d2h"), e2e_dev)},
                    "genotype_upload": ("host pointer, whole matrix" if world == 1 else
                                        "1/%d SNP slab per rank host->device + in-place NCCL all-gather over NVLink" % world),
                    "includes": "genotype matrix upload + mixing, CSR upload, all kernels, result download (ccu_demux_score scores a "
                                "store of this size in two parts, the second part's upload under the first part's kernels; "
                                "device_ms_inside lists the sums of both parts)"
                                + (", all-gather of the result records + copy of the full table to rank 0's host" if world > 1 else ""),
                    "bytes_are": "per rank (rank 0)"},
            "gpu_launches": launches, "roofline": roofline, "roofline_k1": roofline_k1, "kernel_ms": kernel_ms,
            "sample_result": [float(res_check["sngBestLLK"][0]), float(res_check["dblBestLLK"][0])]}


def trim_demux(m, d, world):
    """The part of a measure_demux record an extra.* block carries."""
    r = m["roofline"]
    return {"config": config_dict(d, world), "value": m["value"], "unit": UNIT, "ms_per_step": m["ms_per_step"],
            "e2e": {k: m["e2e"][k] for k in ("value", "ms_per_step", "ms_set_genotypes_call", "ms_score_call",
                                              "ms_result_gather", "h2d_bytes_per_step", "d2h_bytes_per_step")},
            "clocks": m["clocks"], "gpu_launches": m["gpu_launches"],
            "roofline": {k: r[k] for k in ("bound", "kernel", "achieved", "peak", "unit", "frac", "entries_affine",
                                           "entries_general", "k2_ms_avg", "k2_share_of_step",
                                           "algorithmic_speedup_vs_survey_count")},
            "kernel_ms": m["kernel_ms"]}


def measure_fmx(ctx, iters=10):
    """BASELINE configs[3]: freemuxlet EM, K=16, 50k barcodes, 300k SNPs.  Every rank holds the droplet store; the
    E-step is barcode-sharded, the M-step SNP-slab-sharded (cramore_b200/fmx_dist.py).  Timed per exchange form:
    one whole EM run (initial M-step + `iters` iterations + final statistics all-reduce) between CUDA events on the
    shared stream, max over ranks; rank 0 also runs the single-GPU loop and requires identical results."""
    torch, dist = ctx.torch, ctx.dist
    from cramore_b200 import FreemuxEngine
    from cramore_b200.fmx_dist import device_views, posterior_view, run_em_distributed, snp_slabs
    from cramore_b200.shard import balanced_ranges
    from cramore_b200.synth import make_dataset
    world, rank, local, dev = ctx.world, ctx.rank, ctx.local, ctx.dev
    K, GE = 16, 0.0
    d = make_dataset("C4")
    rng = np.random.default_rng(11)
    init = (d["s1"] % K).astype(np.int32)  # stands for --init-cluster (the timed region is the EM loop)
    init[d["is_dbl"]] = -1
    init[rng.random(len(init)) < 0.3] = -1
    nnz = int(d["cell_ptr"][-1])
    eng = FreemuxEngine(local)
    eng.set_stream(ctx.stream.cuda_stream)
    eng.configure_fmx(K, 0.5, GE)
    t0 = time.perf_counter()
    eng.upload_fmx(d["cell_ptr"], d["snp_idx"], d["read_ptr"], d["al"], d["bq"], d["af"])
    t_upload = time.perf_counter() - t0
    setup = eng.fmx_timings()
    stats, assign = device_views(eng, dev)
    post = posterior_view(eng, dev)
    cell_ranges = balanced_ranges(d["cell_ptr"], world)
    slabs = snp_slabs(d["snp_idx"], d["nsnps"], world)

    graphs = {}  # the "push" loop (kernels only, ordered by their own epoch flags) is replayed as one CUDA graph

    def run(exchange):
        ctx.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = eng.fmx_timings()["launches"]
        e0.record()
        tm = {}
        loc, full = run_em_distributed(eng, stats, assign, cell_ranges, slabs, init, niter=iters, post=post,
                                       exchange=exchange, graphs=graphs if exchange == "push" else None, timing=tm)
        e1.record()
        torch.cuda.synchronize()
        run.loop_ms = ctx.max_over_ranks(tm["loop_events"][0].elapsed_time(tm["loop_events"][1]))
        return full, ctx.max_over_ranks(e0.elapsed_time(e1)), eng.fmx_timings()["launches"] - l0

    def phase(fn):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        return ctx.max_over_ranks(a.elapsed_time(b))

    # "push": the fused loop -- M-step kernel forms the posteriors and stores them into every rank's array over NVLink,
    # decision kernel does the same with the assignments, epoch flags inside the kernels order the ranks (at N = 1 the
    # same two launches + decision per iteration without peers); "posteriors": the NCCL form (all-reduce of the
    # posterior array, all-gather of the assignments) kept beside it
    forms = ["push", "posteriors"] if world > 1 else ["push"]
    sampler = ClockSampler(local)
    res = {}
    launches = 0
    sampler.begin()
    for form in forms:
        run(form)  # warm-up: NCCL channels, peer mappings, allocations
        if form == "push":
            l_eager = eng.fmx_timings()["launches"]
            run(form)  # captures the loop into a CUDA graph and replays it once
            l_eager = eng.fmx_timings()["launches"] - l_eager
        full, ms, nl = run(form)
        if form == "push":
            nl = l_eager  # a replay launches the same kernels; the host-side counter only sees the capture
        launches += 2 * nl
        # ms_per_iter: the EM loop alone (initial M-step + `iters` iterations; the initial M-step counted as 0.3 of an
        # iteration) -- the figure comparable across N; ms_whole_run adds what a run pays once: copy of the initial
        # assignment, all-reduce of the statistics array, result download + gather
        res[form] = {"ms_per_iter": run.loop_ms / (iters + 0.3), "ms_loop": run.loop_ms, "ms_whole_run": ms,
                     "launches_per_run": nl, "cuda_graph": form == "push", "_full": full}
    # per-phase device times of this rank's shard (one extra pass, a synchronise around each phase; max over ranks).
    # The fused calls run without their in-kernel waits here: every rank is inside the same phase anyway.
    # (M-step first: it leaves the posteriors of every SNP, on every rank, that the E-step call insists on)
    phases = {"mstep_ms": phase(lambda: eng.mstep_fused_dev(False, False, False))}
    ctx.barrier()
    phases["estep_ms"] = phase(lambda: eng.estep_fused_dev(False))
    ctx.barrier()
    phases["mstep_with_stats_records_ms"] = phase(lambda: eng.mstep_fused_dev(False, True, False))
    if world > 1:
        phases["rank_barrier_kernel_ms"] = phase(lambda: eng.rank_barrier_dev())
        phases["allreduce_posteriors_ms"] = phase(lambda: dist.all_reduce(post))
    sampler.end()
    clocks = sampler.summary()

    ok, single_ms = None, None
    if rank == 0:
        if world > 1:
            ref_eng = FreemuxEngine(local)
            ref_eng.configure_fmx(K, 0.5, GE)
            ref_eng.upload_fmx(d["cell_ptr"], d["snp_idx"], d["read_ptr"], d["al"], d["bq"], d["af"])
            ref_eng.run_em(init, niter=iters)
            t0 = time.perf_counter()
            ref, _ = ref_eng.run_em(init, niter=iters)
            single_ms = 1e3 * (time.perf_counter() - t0) / iters
            ok = all(bool(np.array_equal(ref, res[f]["_full"])) for f in forms)
            ref_eng.close()
    ctx.barrier()
    full = res[forms[0]]["_full"]
    out = None
    if rank == 0:
        try:
            hbm_peak, peak_src = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "MEASURED_PEAKS.json"
        except Exception:
            hbm_peak, peak_src = 6650.0, "fallback of B200_PROFILING.md"
        nobs = int(np.unique(d["snp_idx"]).size)
        # Algorithmic bytes per launch, per-rank shard = 1/N of each.  E-step: a single-read entry is
        # 20 B ((gl[0]/2, b) + SNP id) + 8*K B of gathered mean genotypes, an entry with >= 2 reads 84 B + 24*K B
        # (SURVEY 8(d)'s figure).  M-step: 88 B pileup + 4 B droplet id per entry in, (24 + 8) B of posteriors and mean
        # genotypes per (cluster, SNP of the slab) out (the 96-byte statistics records only on the last iteration).
        al_ne2 = (np.asarray(d["al"]) != 2).astype(np.int64)
        csum = np.concatenate([[0], np.cumsum(al_ne2)])
        neff = csum[np.asarray(d["read_ptr"][1:])] - csum[np.asarray(d["read_ptr"][:-1])]
        n_gen = int((neff > 1).sum())
        n_aff = nnz - n_gen
        e_bytes = ((20 + 8 * K) * n_aff + (84 + 24 * K) * n_gen) / world
        m_bytes = (92 * nnz + 32 * K * d["nsnps"]) / world
        # FP64 instructions of the E-step as built, per entry
        e_flops = 2.0 * (fmx_estep_fp64_instr(K, False) * n_aff + fmx_estep_fp64_instr(K, True) * n_gen) / world
        fp64_peak = ctx.fp64_peak
        roof = [{"kernel": "fmx E-step (F2)", "bound": "hbm (SURVEY 8d)", "achieved": e_bytes / (phases["estep_ms"] * 1e-3) / 1e9,
                 "peak": hbm_peak, "unit": "GB/s", "frac": e_bytes / (phases["estep_ms"] * 1e-3) / 1e9 / hbm_peak,
                 "algorithmic_bytes_per_launch": int(e_bytes), "ms": phases["estep_ms"], "peak_source": peak_src,
                 "fp64": {"achieved": e_flops / (phases["estep_ms"] * 1e-3) / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
                          "frac": e_flops / (phases["estep_ms"] * 1e-3) / 1e12 / fp64_peak if fp64_peak else None,
                          "note": "needed FP64 flops (2 per pair and SNP on single-read entries) / time; the mean-genotype "
                                  "array (K*8 B per SNP) lives in L2, so HBM does not bind"},
                 "binding": {"resource": "instruction issue (warp schedulers) and the L2 round trip of the operand fetch one step ahead",
                             "frac_ncu": 0.59,
                             "source": "profiles/r02b_fmx_kernels_ncu_full.csv: smsp__issue_active 59.5 %, FP64 pipe 43 %, "
                                       "shared-memory wavefronts 45 %, DRAM 4 % of peak at C4 on one GPU (static figures "
                                       "of that capture, not measured in this run)"}},
                {"kernel": "fmx M-step (F3)", "bound": "hbm", "achieved": m_bytes / (phases["mstep_ms"] * 1e-3) / 1e9,
                 "peak": hbm_peak, "unit": "GB/s", "frac": m_bytes / (phases["mstep_ms"] * 1e-3) / 1e9 / hbm_peak,
                 "algorithmic_bytes_per_launch": int(m_bytes), "ms": phases["mstep_ms"], "peak_source": peak_src,
                 "binding": {"resource": "latency of the ordered fold (dependent chain per warp) at 28 warps per SM",
                             "frac_ncu": 0.53,
                             "source": "profiles/r02b_fmx_kernels_ncu_full.csv: smsp__issue_active 52.7 %, FP64 pipe 28 %, "
                                       "DRAM 46 % of peak at C4 on one GPU (static figures of that capture)"}}]
        out = {"config": {"workload": "C4 freemuxlet EM: K=%d, %d barcodes, %d SNPs, %d entries, %d iterations; E-step "
                                      "barcode-sharded, M-step SNP-slab-sharded over %d GPU(s)" %
                                      (K, d["nbcs"], d["nsnps"], nnz, iters, world)},
               "metric": "ms per EM iteration (EM loop between device events / (iterations + 0.3 for the initial M-step), max over ranks)",
               "scaling": "strong", "n_gpus": world, "iters": iters,
               "exchange": {f: {k: v for k, v in res[f].items() if not k.startswith("_")} for f in forms},
               "ms_per_iter": res[forms[0]]["ms_per_iter"], "phase_ms": phases, "clocks": clocks,
               "entries_single_read": n_aff, "entries_general": n_gen,
               "matches_single_gpu": ok, "single_gpu_ms_per_iter_wall": single_ms,
               "speedup_vs_single_gpu_wall": (single_ms / res[forms[0]]["ms_per_iter"]) if single_ms else None,
               "posterior_exchange_bytes": int(post.numel() * 8), "stats_allreduce_bytes_once": int(stats.numel() * 8),
               # push form, per rank and iteration: its slab's mean genotypes (8 B per SNP and cluster) to every peer, and
               # posterior triples only at SNPs with an entry of >= 2 reads (at most one SNP per such entry)
               "push_bytes_per_rank_per_iter": int((world - 1) / world * K * (8 * d["nsnps"] + 24 * min(n_gen, d["nsnps"])))
               if K <= 16 else int((world - 1) / world * K * 32 * d["nsnps"]),
               "setup": setup, "upload_s": t_upload, "roofline_fmx": roof, "gpu_launches": launches,
               "singlets": int((full["type"] == 0).sum()), "doublets": int((full["type"] == 1).sum())}
    graphs.clear()
    torch.cuda.synchronize()
    eng.close()
    return out


def fmx_estep_fp64_instr(K, general):
    """FP64 instructions per (droplet, SNP) entry the algorithm needs: an entry with >= 2 reads takes
    t = GL^T p per cluster (9) and 4 per pair; a single-read entry u = a/2 + b m per cluster (1) and 2 per pair."""
    npairs = K * (K + 1) // 2
    return (4 * npairs + 9 * K) if general else (2 * npairs + K)


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    from cramore_b200.synth import make_dataset
    ctx = Ctx()
    world, rank = ctx.world, ctx.rank

    # ---- headline: the synthetic C2 workload, same genotypes on every rank, rank-specific droplets (weak scaling)
    d = make_dataset(args.workload, nbcs=args.nbcs, droplet_seed_offset=rank)
    m = measure_demux(ctx, d, args.steps, args.warmup)
    ctx.fp64_peak = m["fp64_peak"]
    line = None
    if rank == 0:
        cpu_val, cpu_n, cpu_dt, cpu_kind = cpu_rate(d, args.cpu_seconds, 1)
        line = {"metric": METRIC, "value": m["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": m["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config_dict(d, world), "clocks": m["clocks"], "e2e": m["e2e"],
                "gpu_launches": m["gpu_launches"], "roofline": m["roofline"], "roofline_k1": m["roofline_k1"],
                "cpu_baseline": {"value": cpu_val, "unit": UNIT, "cores": 1, "kind": cpu_kind,
                                 "sample": "first %d barcodes of the same workload, %.1f s, single thread "
                                           "(the reference is single-threaded); kind=reference is "
                                           "oracle/_ref/ref_cramore, the reference's own engine sources compiled "
                                           "behind a zlib I/O shim" % (cpu_n, cpu_dt)},
                "kernel_ms": m["kernel_ms"], "sample_result": m["sample_result"], "extra": {}}
    del d

    # ---- the other named configs, in the same process so that they sit on the driver's record.  A watchdog prints
    # the headline line alone if they do not finish (every rank leaves, so nobody is left inside a collective).
    done = threading.Event()

    def emit_and_leave(why):
        if rank == 0:
            line["extra"]["_aborted"] = why
            print(json.dumps(line), flush=True)
        os._exit(0)

    if not args.no_extras:
        wd = threading.Timer(args.extras_deadline, lambda: (not done.is_set()) and emit_and_leave("extras deadline reached"))
        wd.daemon = True
        wd.start()
        want = [x for x in args.extras.split(",") if x]
        t_extras = time.perf_counter()

        def run_extra(name, fn):
            """fn() -> record (rank 0) or None.  Exceptions are caught per rank; the ranks then agree on success so
            that nobody enters the next block's collectives alone."""
            t0 = time.perf_counter()
            rec, err = None, None
            try:
                rec = fn()
            except Exception as e:  # noqa: BLE001
                err = "%s: %s" % (type(e).__name__, str(e)[:300])
            ok = ctx.min_over_ranks(0.0 if err else 1.0)
            if rank == 0:
                if ok < 1.0:
                    line["extra"][name] = {"error": err or "failed on another rank"}
                else:
                    rec["seconds"] = time.perf_counter() - t0
                    line["extra"][name] = rec
            return ok >= 1.0

        def demux_extra(name, nbcs, steps, **kw):
            def fn():
                dd = make_dataset(name, nbcs=nbcs, droplet_seed_offset=rank, **kw)
                mm = measure_demux(ctx, dd, steps, 3, fp64_peak=ctx.fp64_peak)
                rec = trim_demux(mm, dd, world)
                rec["steps"] = steps
                return rec
            return fn

        if "c3" in want:
            # BASELINE configs[2] / north_star target: 100k barcodes over 8 GPUs = 12.5k per GPU
            if run_extra("c3", demux_extra("C3", 12500, 5)) and rank == 0:
                c3 = line["extra"]["c3"]
                tot = 12500 * world
                c3["target"] = {"north_star": ">= 100k droplets x 32,640 pairs x 11 alphas x 500k SNPs in < 60 s on 8 GPUs",
                                "droplets_this_run": tot,
                                "seconds_per_pass_e2e": c3["e2e"]["ms_per_step"] * 1e-3,
                                "seconds_for_100k_droplets_at_this_rate": 100000.0 / c3["e2e"]["value"],
                                "is_the_full_target_config": world == 8}
        if "c5" in want:
            run_extra("c5", demux_extra("C5", None, 3))
        if "c2_skew" in want:
            # item "general path": a Zipf SNP draw puts >= 30 % of the entries on the >= 2-reads-per-SNP path of K2
            def skew():
                dd = make_dataset("C2", nbcs=5000, droplet_seed_offset=rank, snp_skew=1.5, rbar=800)
                mm = measure_demux(ctx, dd, 3, 3, fp64_peak=ctx.fp64_peak)
                rec = trim_demux(mm, dd, world)
                rec["config"]["snp_skew"] = "SNP of a read ~ Zipf(s=1.5) over a random permutation of the panel, 800 reads per cell"
                rec["steps"] = 3
                return rec
            run_extra("c2_skew", skew)
        if "fmx_c4" in want:
            run_extra("fmx_c4", lambda: measure_fmx(ctx))
        if rank == 0:
            line["extra"]["_seconds"] = time.perf_counter() - t_extras
            line["gpu_launches"] += sum(v.get("gpu_launches", 0) for v in line["extra"].values() if isinstance(v, dict))
        done.set()
        wd.cancel()

    if world > 1:
        ctx.dist.barrier()
        ctx.dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
