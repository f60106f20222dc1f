"""Development helper for compute-sanitizer runs: TMA stage kernel vs direct kernel on a small ragged lattice.
    python scripts/sanity_tile.py '<json options>' <T>"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def one(opts, T):
    import numpy as np
    from jams_b200 import workloads as W
    w = W.c3_sc(dims=(12, 20, 70), temperature=T)
    res = []
    for o in (dict(kernel=0), opts):
        s = W.make_solver(w, options=o, random_spins_seed=3, seed=11)
        s.run(7)
        res.append(s.spins())
    print("opts", opts, "T", T, "max|tile-direct| =", float(np.abs(res[0] - res[1]).max()), flush=True)


if __name__ == "__main__":
    one(json.loads(sys.argv[1]), float(sys.argv[2]))
