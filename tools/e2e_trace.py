"""Host wall-clock trace of the end-to-end demuxlet call (ccu_demux_set_genotypes + ccu_demux_score with pinned host
buffers) on the C2 workload: CCU_TRACE=1 makes the library print the phases of every score call to stderr."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["CCU_TRACE"] = "1"


def main():
    import torch
    from cramore_b200 import RESULT_DTYPE, DemuxEngine
    from cramore_b200.synth import make_dataset
    d = make_dataset(sys.argv[1] if len(sys.argv) > 1 else "C2")
    host, keep = {}, []
    for k in ("gp", "snp_err", "cell_ptr", "snp_idx", "read_ptr", "al", "bq"):
        t = torch.from_numpy(np.ascontiguousarray(d[k])).pin_memory()
        keep.append(t)
        host[k] = t.numpy()
    out_t = torch.empty(d["nbcs"] * RESULT_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
    out = out_t.numpy().view(RESULT_DTYPE)
    eng = DemuxEngine(0)
    eng.configure(d["nv"], d["alphas"], 0.5)
    for i in range(6):
        t0 = time.perf_counter()
        eng.set_genotypes(host["gp"], host["snp_err"])
        t1 = time.perf_counter()
        eng.score(host["cell_ptr"], host["snp_idx"], host["read_ptr"], host["al"], host["bq"], out)
        t2 = time.perf_counter()
        print("step %d: set_genotypes %.3f ms, score %.3f ms; engine timings %s" % (
            i, 1e3 * (t1 - t0), 1e3 * (t2 - t1), {k: round(v, 3) for k, v in eng.timings().items() if k.startswith("ms_")}),
            file=sys.stderr, flush=True)


if __name__ == "__main__":
    main()
