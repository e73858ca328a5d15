"""Writes tests/golden/appendix_c.json: the hand-derived known-answer vector of SURVEY.md App. C
(derived from cmd_cram_demuxlet.cpp:655-821; the reference ships no fixtures).  The expected
numbers below are the ones printed in SURVEY.md, NOT produced by this repository's oracle."""
import json
import os

G = {
    "alpha": [0.0, 0.5], "doublet_prior": 0.5,
    "gps": [[[0.9, 0.08, 0.02], [0.05, 0.15, 0.80]], [[0.7, 0.2, 0.1], [0.1, 0.3, 0.6]]],
    "cell_ptr": [0, 2], "snp_idx": [0, 1], "read_ptr": [0, 2, 3], "al": [0, 0, 1], "bq": [20, 20, 20],
    "tiles": [
        [1, 1, 1] + [0.25168633593625139] * 3 + [1.1336811672278334e-05] * 3 +
        [1, 0.56376333485072949, 0.25168633593625139, 0.56376333485072949, 0.25168633593625139,
         0.063769003256565662, 0.25168633593625139, 0.063769003256565662, 1.1336811672278334e-05],
        [0.0033670034666666675] * 3 + [0.50168350173333331] * 3 + [1, 1, 1] +
        [0.0033670034666666675, 0.25252525260000003, 0.50168350173333331, 0.25252525260000003,
         0.50168350173333331, 0.7508417508666666, 0.50168350173333331, 0.7508417508666666, 1],
    ],
    "llk": [[[-1.6792945214601431, -1.6991368151076829], [-1.6792945214601431, -1.9050874441560064]],
            [[-2.7196868150951596, -1.9050874441560066], [-2.7196868150951601, -3.2418012776880158]]],
    "sumLLK": 0.1288373010523936, "sngLLK": 0.061189378524528965, "sngPP": 0.934589463554648,
    "type": "AMB", "sBest": 0, "sNext": 1, "dBest": [0, 1, 1], "dNext": [1, 0, 1],
}
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "appendix_c.json"), "w") as f:
    json.dump(G, f, indent=1)
