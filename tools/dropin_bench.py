"""The drop-in, timed: the SAME command line on the SAME files through the stock reference (oracle/_ref/ref_cramore:
cmdCramDemuxlet compiled unmodified) and through the patched reference (oracle/_ref/ref_cramore_gpu: the same
function with its likelihood loop replaced by calls into libcramore_cuda.so).  Wall clock of
the whole process -- option parsing, load_from_plp, the VCF join and .best writing are the reference's own code in
both.  C2-shaped droplets (64 samples, 11 alphas, ~200 reads per droplet) on a reduced number of barcodes and SNPs so
that the single-threaded reference finishes in about half a minute.  Prints one JSON line."""
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from cramore_b200.synth import make_dataset, write_plp, write_vcf
    nbcs = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    d = make_dataset("C2", nbcs=nbcs, nsnps=50000, with_barcodes=True)
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        write_plp(d, os.path.join(tmp, "in"))
        write_vcf(d, os.path.join(tmp, "in.vcf"))
        args = ["demuxlet", "--plp", "in", "--vcf", "in.vcf", "--field", "GP"]
        for a in d["alphas"]:
            args += ["--alpha", repr(float(a))]
        for tag, exe in (("gpu", "ref_cramore_gpu"), ("cpu", "ref_cramore")):
            t0 = time.perf_counter()
            r = subprocess.run([os.path.join(ROOT, "oracle", "_ref", exe)] + args + ["--out", tag], cwd=tmp, capture_output=True, text=True)
            out[tag + "_s"] = time.perf_counter() - t0
            if r.returncode != 0:
                raise SystemExit(r.stderr[-2000:])
        a = open(os.path.join(tmp, "gpu.best")).read().splitlines()
        b = open(os.path.join(tmp, "cpu.best")).read().splitlines()
        same = sum(x == y for x, y in zip(a, b))
    print(json.dumps({"what": "reference command line, stock vs patched onto libcramore_cuda.so, whole-process wall clock",
                      "config": {"workload": "C2-shaped: 64 samples, 11 alphas, ~200 reads/droplet, %d barcodes, 50000 SNPs" % nbcs},
                      "stock_reference_s": out["cpu_s"], "patched_reference_s": out["gpu_s"],
                      "speedup": out["cpu_s"] / out["gpu_s"], "rows": len(b) - 1, "rows_identical_text": same - 1}), flush=True)


if __name__ == "__main__":
    main()
