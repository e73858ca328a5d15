"""Small fixed demuxlet run for ncu captures: a BASELINE shape, a few thousand barcodes, 3 engine runs.
usage: python tools/prof_demux.py [nbcs] [config] [snp_skew] [rbar]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cramore_b200 import DemuxEngine
from cramore_b200.synth import make_dataset

nbcs = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
name = sys.argv[2] if len(sys.argv) > 2 else "C2"
skew = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
rbar = int(sys.argv[4]) if len(sys.argv) > 4 else None
d = make_dataset(name, nbcs=nbcs, snp_skew=skew, rbar=rbar)
eng = DemuxEngine(0)
eng.configure(d["nv"], d["alphas"], 0.5)
eng.set_genotypes(d["gp"], d["snp_err"])
eng.upload(d["cell_ptr"], d["snp_idx"], d["read_ptr"], d["al"], d["bq"])
for _ in range(3):
    eng.run()
    eng.wait()
print(eng.timings())
eng.close()
