"""Development helper for compute-sanitizer runs: TMA stage kernel vs direct kernel on a small ragged lattice.
    python scripts/sanity_tile.py '<json options>' <T>"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def one(opts, T, which="sc"):
    import numpy as np
    from jams_b200 import workloads as W
    from jams_b200.lattice import Lattice, Material
    if which == "c4":     # BASELINE config 4's template (bcc, 8 shells, reach 2) on a small ragged lattice with one open axis
        w = W.c4_bcc_long_range(8, temperature=T)
        w["lattice"] = Lattice([Material("Fe", 2.2, alpha=0.1)], np.eye(3), [("Fe", (0, 0, 0)), ("Fe", (0.5, 0.5, 0.5))], (7, 9, 37), periodic=(True, True, False))
    elif which == "bcc":
        w = W.c2_bcc_fe(8, temperature=T)
        w["lattice"] = Lattice([Material("Fe", 2.2, alpha=0.1)], np.eye(3), [("Fe", (0, 0, 0)), ("Fe", (0.5, 0.5, 0.5))], (9, 7, 66))
    else:
        w = W.c3_sc(dims=(12, 20, 70), temperature=T)
    res = []
    for o in (dict(kernel=0), opts):
        s = W.make_solver(w, options=o, random_spins_seed=3, seed=11)
        s.run(7)
        res.append(s.spins())
    print("opts", opts, "T", T, "max|tile-direct| =", float(np.abs(res[0] - res[1]).max()), flush=True)


if __name__ == "__main__":
    one(json.loads(sys.argv[1]), float(sys.argv[2]), sys.argv[3] if len(sys.argv) > 3 else "sc")
