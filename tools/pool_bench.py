import sys
sys.path.insert(0, '/root/repo')
import numpy as np
from cramore_b200 import DemuxEngine
from cramore_b200.synth import make_dataset
eng = DemuxEngine(0)
for nv, alphas in ((8, [0, 0.5]), (16, [0, 0.5]), (16, [round(0.05*i,2) for i in range(11)]), (12, [0,0.25,0.5]), (32, [0, 0.5]), (24, [round(0.05*i,2) for i in range(11)])):
    d = make_dataset("C2", nbcs=20000, nv=nv)
    d["alphas"] = np.asarray(alphas, np.float64)
    eng.configure(d["nv"], d["alphas"], 0.5)
    eng.set_genotypes(d["gp"], d["snp_err"])
    eng.upload(d["cell_ptr"], d["snp_idx"], d["read_ptr"], d["al"], d["bq"])
    for _ in range(3):
        eng.run(); eng.wait()
    t = eng.timings()
    print(nv, len(alphas), "ms_score %.3f ms_singlet %.3f ms_run %.3f" % (t["ms_score"], t["ms_singlet"], t["ms_run"]), "ctas", t["score_ctas"])
eng.close()
