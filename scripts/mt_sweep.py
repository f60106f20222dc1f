"""BASELINE config 3 as the reference would run it: an M(T) sweep on simple cubic NN Heisenberg + Zeeman, one chained run
(the state of temperature k is the start of temperature k + 1, README.md "temperature sweep" idiom of JAMS: physics.temperature
patched per run, lattice.spins = the previous run's final state).  Everything stays on the device: jb_step with the new
temperature, the magnetisation monitor's reduction every `every` steps.

    python scripts/mt_sweep.py [--n 256] [--temps 50,150,250,330,400,500] [--equil 4000] [--meas 4000] [--every 20] [--alpha 0.5] [--dt 5e-16]

Prints one line per temperature (<m_z>, <|m|>, their standard errors, ms per step) and a JSON record."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from jams_b200.lattice import Lattice, Material
from jams_b200.solver import create_hamiltonian, create_solver


def sweep(n, temps, equil, meas, every, alpha=0.5, dt=5e-16, seed=11, dims=None, log=print):
    dims = dims or (n, n, n)
    lat = Lattice([Material("A", 2.0, alpha=alpha)], np.eye(3), [("A", (0, 0, 0))], dims, gilbert_prefactor=True)
    hams = [dict(module="exchange", interactions=[("A", "A", [1.0, 0.0, 0.0], 3.5e-21)]), dict(module="zeeman", dc_local_field=[[0.0, 0.0, 1.0]])]
    s = create_solver(dict(module="llg-heun-b200-gpu", t_step=dt, t_max=1.0, seed=seed, gilbert_prefactor=True), lat)
    for h in hams:
        s.register_hamiltonian(create_hamiltonian(h, lat))
    s.set_spins(np.tile([0.0, 0.0, 1.0], (lat.num_spins, 1)))
    mu_total = None
    out = []
    for T in temps:
        s.set_temperature(T)
        s.run(equil)
        mz, mabs = [], []
        s.ctx.synchronize()
        t0 = time.perf_counter()
        for _ in range(meas // every):
            s.run(every)
            m4 = s.ctx.magnetisation(None, 1)[0]   # sum mu s (3) and sum mu
            mu_total = m4[3]
            m = m4[:3] / mu_total
            mz.append(m[2]); mabs.append(float(np.linalg.norm(m)))
        s.ctx.synchronize()
        ms = (time.perf_counter() - t0) * 1e3 / meas
        mz, mabs = np.array(mz), np.array(mabs)
        blocks = max(4, len(mz) // 10)   # standard error from block means (the samples are correlated)
        bm = np.array([b.mean() for b in np.array_split(mz, blocks)])
        rec = dict(T=T, mz=float(mz.mean()), mz_err=float(bm.std(ddof=1) / np.sqrt(blocks)), m_abs=float(mabs.mean()), ms_per_step=ms)
        out.append(rec)
        log(f"T = {T:7.1f} K   <m_z> = {rec['mz']:.5f} +- {rec['mz_err']:.5f}   <|m|> = {rec['m_abs']:.5f}   {ms:.4f} ms/step (monitor every {every})")
    return lat, out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--temps", default="50,150,250,330,400,500")
    ap.add_argument("--equil", type=int, default=4000)
    ap.add_argument("--meas", type=int, default=4000)
    ap.add_argument("--every", type=int, default=20)
    ap.add_argument("--alpha", type=float, default=0.5)
    ap.add_argument("--dt", type=float, default=5e-16)
    a = ap.parse_args()
    temps = [float(t) for t in a.temps.split(",")]
    lat, out = sweep(a.n, temps, a.equil, a.meas, a.every, a.alpha, a.dt)
    print(json.dumps(dict(workload=f"C3 sc {a.n}^3 NN Heisenberg + Zeeman (1 T), M(T) sweep, chained temperatures", spins=lat.num_spins, alpha=a.alpha, dt_s=a.dt,
                          equil_steps=a.equil, meas_steps=a.meas, sweep=out)))
