"""Development helper: tile kernel vs direct kernel on a small lattice (each option set in its own process)."""
import json, os, subprocess, sys
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
    if len(sys.argv) > 1:
        one(json.loads(sys.argv[1]), float(sys.argv[2]))
        sys.exit(0)
    for opts in (dict(kernel=3, row_offset=4), dict(kernel=3, row_offset=8), dict(kernel=2, row_offset=4), dict(kernel=2, row_offset=8), dict(kernel=1, row_offset=8), dict(kernel=3), dict(kernel=3, tile_y=4, tile_z=32), dict(kernel=3, tile_y=3, tile_z=8, ring=4), dict(kernel=2), dict(kernel=2, spt=2), dict(kernel=2, tile_y=3, tile_z=8, ring=4), dict(kernel=2, spt=2, tile_y=8, tile_z=32, ring=9, ring_u=4), dict(kernel=1)):
        for T in (0.0, 50.0):
            r = subprocess.run([sys.executable, os.path.abspath(__file__), json.dumps(opts), str(T)], capture_output=True, text=True, timeout=300)
            print((r.stdout or "").strip() or f"opts {opts} T {T} CRASHED: {(r.stderr or '').strip()[-400:]}", flush=True)
