#!/usr/bin/env python
"""Per-phase device times of the fused freemuxlet loop at C4 on one GPU (A/B of kernel variants: CCU_LIB selects the
library).  usage: [CCU_LIB=...] python tools/fmx_phases.py [reps] [snp_skew]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cramore_b200 import LIB_PATH, FreemuxEngine
from cramore_b200.synth import make_dataset

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
skew = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
K = 16
d = make_dataset("C4", snp_skew=skew) if skew > 0 else make_dataset("C4")
init = (d["s1"] % K).astype(np.int32)
init[d["is_dbl"]] = -1
torch.cuda.set_device(0)
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
eng = FreemuxEngine(0)
eng.set_stream(st.cuda_stream)
eng.configure_fmx(K, 0.5, 0.0)
eng.upload_fmx(d["cell_ptr"], d["snp_idx"], d["read_ptr"], d["al"], d["bq"], d["af"])
res, _ = eng.run_em(init, niter=3)


def timed(fn):
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return {"min": min(ts), "avg": float(np.mean(ts))}


out = {"lib": os.path.basename(LIB_PATH), "snp_skew": skew, "entries": int(d["cell_ptr"][-1]),
       "estep": timed(lambda: eng.estep_fused_dev(False)),
       "mstep": timed(lambda: eng.mstep_fused_dev(False, False, False)),
       "mstep_stats": timed(lambda: eng.mstep_fused_dev(False, True, False))}
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
a.record()
for _ in range(20):
    eng.estep_fused_dev(False)
    eng.mstep_fused_dev(False, False, False)
b.record()
torch.cuda.synchronize()
out["iter_ms"] = a.elapsed_time(b) / 20
out["singlets"] = int((res["type"] == 0).sum())
print(json.dumps(out))
eng.close()
