"""Multi-process check of the slab decomposition (one rank per GPU, NCCL plumbing, P2P halo stores):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/mgpu_check.py
Every rank steps its x-slab; rank 0 gathers the spins and compares them bit for bit with an undecomposed run of the
same lattice on its own GPU (the Langevin noise is keyed by the global site id, so the trajectories must be identical)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from jams_b200 import workloads as W
from jams_b200.distributed import TorchComm
from jams_b200.solver import MagnetisationMonitor


def pinned_boundaries_case(rank, world, local):
    """pinned_boundaries on a slab-decomposed open-x wall: jb_rotate_region stores rotated ghost images into the neighbours'
    boxes and must be ordered by the same epoch handshake as the stages; front / back regions span every slab face"""
    from jams_b200.solver import PinnedBoundariesPhysics
    w = W.c1_bloch_wall((16 * world, 8, 12))
    lat = w["lattice"]
    settings = dict(module="pinned_boundaries", left_pinned_magnetisation=[0, 0, 1.0], right_pinned_magnetisation=[0, 0, -1.0],
                    front_pinned_magnetisation=[0, 1.0, 0], back_pinned_cells=2, back_pinned_magnetisation=[1.0, 0, 0])
    comm = TorchComm(periodic_x=False, device=f"cuda:{local}")
    s = W.make_solver(w, comm=comm, seed=3, device=local)
    s0 = lat.initial_spins(seed=9)
    per = lat.num_spins // world
    s.set_spins(s0[rank * per:(rank + 1) * per])
    s.register_physics_module(PinnedBoundariesPhysics(settings, lat))
    for _ in range(8):
        s.update_physics_module()
        s.run(1)
    mine = torch.from_numpy(s.spins()).to(f"cuda:{local}")
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    comm.barrier(s.ctx)
    ok = True
    if rank == 0:
        got = torch.cat(parts).cpu().numpy()
        single = W.make_solver(w, seed=3, device=local)
        single.set_spins(s0)
        single.register_physics_module(PinnedBoundariesPhysics(settings, lat))
        for _ in range(8):
            single.update_physics_module()
            single.run(1)
        diff = float(np.abs(single.spins() - got).max())
        ok = diff <= 1e-13   # the all-reduced region moment sums the slabs' partial sums in a different order
        print(f"mgpu_check[{world} ranks] pinned_boundaries (left/right/front/back) on an open-x wall: max diff = {diff:.3e}, ok = {ok}", flush=True)
        single.ctx.close()
    comm.barrier(s.ctx)
    s.ctx.close()
    return ok


def alloy(world, temperature):
    """lattice.impurities: 30 % Co on the Fe sites of a bcc lattice -- not translation invariant, so the exchange travels as the
    general neighbour list (GLOBAL site ids on every rank, neighbours across a slab face addressed through the x ghost planes)"""
    from jams_b200.lattice import Lattice, Material
    lat = Lattice([Material("Fe", 2.2, alpha=0.1), Material("Co", 1.7, alpha=0.05)], np.eye(3), [("Fe", (0, 0, 0)), ("Fe", (0.5, 0.5, 0.5))],
                  (6 * world, 7, 9), impurities=[("Fe", "Co", 0.3)], impurities_seed=5)
    hams = [dict(module="exchange", interactions=[("Fe", "Fe", [0.5, 0.5, 0.5], 3.2e-21), ("Fe", "Fe", [1.0, 0.0, 0.0], 1.6e-21)]),
            dict(module="uniaxial", order="K1", anisotropies=[("Fe", [0.0, 0.0, 1.0], 1e-23), ("Co", [1.0, 0.0, 0.0], 4e-23)])]
    return dict(name="alloy", lattice=lat, hamiltonians=hams, spins=None, temperature=temperature)


def deep(world, temperature):
    """BASELINE config 4's template (bcc, eight shells, reach 2) on a small lattice: the rows kernel over the slabs"""
    from jams_b200.lattice import Lattice, Material
    w = W.c4_bcc_long_range(8, temperature=temperature)
    w["lattice"] = Lattice([Material("Fe", 2.2, alpha=0.1)], np.eye(3), [("Fe", (0, 0, 0)), ("Fe", (0.5, 0.5, 0.5))], (8 * world, 9, 40))
    return w


def solver_of(w, comm, seed, device, options):
    """options may carry "solver": "rk4" (the RK4 stages on the TMA ring, four exchanges per step)"""
    options = dict(options or {})
    if options.pop("solver", "heun") == "rk4":
        from jams_b200.solver import create_hamiltonian, create_solver
        lat = w["lattice"]
        s = create_solver(dict(module="llg-rk4-b200-gpu", t_step=W.T_STEP, t_max=1e-9, seed=seed, options=options, device=device), lat, comm)
        for h in w["hamiltonians"]:
            s.register_hamiltonian(create_hamiltonian(h, lat))
        s.set_temperature(w.get("temperature", 0.0))
        return s
    return W.make_solver(w, comm=comm, seed=seed, device=device, options=options or None)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    # default options: the recover_u data flow with the epoch handshake folded into the stage kernels (two launches per step);
    # the other cases force the stored-u data flow, the separate wait / signal launches, and a lattice large enough for several
    # x-chunks, a taper and more work items than resident CTAs per slab
    for name, make, periodic_x, T, steps, opts in (("bcc NN+NNN periodic T=50", lambda: W.c2_bcc_fe(8 * world, temperature=50.0), True, 50.0, 12, None),
                                                   ("sc open-x wall T=0", lambda: W.c1_bloch_wall((16 * world, 8, 40)), False, 0.0, 15, None),
                                                   ("sc periodic T=0", lambda: W.c3_sc(dims=(12 * world, 10, 36), temperature=0.0), True, 0.0, 15, None),
                                                   ("sc periodic T=30 stored u", lambda: W.c3_sc(dims=(12 * world, 10, 36), temperature=30.0), True, 30.0, 15, dict(recover_u=0)),
                                                   ("sc periodic T=30 separate wait/signal launches", lambda: W.c3_sc(dims=(12 * world, 10, 36), temperature=30.0), True, 30.0, 15, dict(fold_halo=0)),
                                                   ("bcc random alloy (lattice.impurities), general neighbour list, T=40", lambda: alloy(world, 40.0), True, 40.0, 12, None),
                                                   ("bcc 8 shells (112 neighbours per spin), rows kernel, T=60", lambda: deep(world, 60.0), True, 60.0, 6, None),
                                                   ("sc periodic T=30, RK4 on the ring", lambda: W.c3_sc(dims=(12 * world, 10, 36), temperature=30.0), True, 30.0, 8, dict(solver="rk4")),
                                                   ("sc 64/rank x 96 x 256 periodic T=80, many items", lambda: W.c3_sc(dims=(64 * world, 96, 256), temperature=80.0), True, 80.0, 10, None),
                                                   ("sc 64/rank x 96 x 256 periodic T=80, short chunks", lambda: W.c3_sc(dims=(64 * world, 96, 256), temperature=80.0), True, 80.0, 10, dict(chunk_long=6, chunk_short=2, tail_pct=50))):
        w = make()
        lat = w["lattice"]
        comm = TorchComm(periodic_x=periodic_x, device=f"cuda:{local}")
        s = solver_of(w, comm, 77, local, opts)
        s0 = w["spins"] if w.get("spins") is not None else lat.initial_spins(seed=5)
        per = lat.num_spins // world
        s.set_spins(s0[rank * per:(rank + 1) * per])
        s.run(steps)
        mine = torch.from_numpy(s.spins()).to(f"cuda:{local}")
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        mag = MagnetisationMonitor(dict(grouping="none"), lat).update(s)
        comm.barrier(s.ctx)
        if rank == 0:
            got = torch.cat(parts).cpu().numpy()
            single = solver_of(w, None, 77, local, opts)
            single.set_spins(s0)
            single.run(steps)
            want = single.spins()
            mag1 = MagnetisationMonitor(dict(grouping="none"), lat).update(single)
            same = bool(np.array_equal(got, want))
            mag_ok = bool(np.allclose(mag[2:], mag1[2:], rtol=0, atol=1e-13))
            print(f"mgpu_check[{world} ranks] {name}: trajectories identical = {same}, max diff = {np.abs(got - want).max():.3e}, "
                  f"all-reduced magnetisation ok = {mag_ok}", flush=True)
            ok = ok and same and mag_ok
            single.ctx.close()
        comm.barrier(s.ctx)
        s.ctx.close()
    ok = pinned_boundaries_case(rank, world, local) and ok
    flag = torch.tensor([1 if ok else 0], device=f"cuda:{local}")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    if rank == 0:
        print("MGPU_CHECK_OK" if ok else "MGPU_CHECK_FAILED", flush=True)
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
