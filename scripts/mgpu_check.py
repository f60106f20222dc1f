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


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    # the T = 0 cases run the pair kernel's recover_u data flow (no stored Heun intermediate, corrector in place); the last case forces it at T > 0
    for name, make, periodic_x, T, steps, opts in (("bcc NN+NNN periodic T=50", lambda: W.c2_bcc_fe(8 * world, temperature=50.0), True, 50.0, 12, None),
                                                   ("sc open-x wall T=0", lambda: W.c1_bloch_wall((16 * world, 8, 40)), False, 0.0, 15, None),
                                                   ("sc periodic T=0", lambda: W.c3_sc(dims=(12 * world, 10, 36), temperature=0.0), True, 0.0, 15, None),
                                                   ("sc periodic T=30 recover_u", lambda: W.c3_sc(dims=(12 * world, 10, 36), temperature=30.0), True, 30.0, 15, dict(recover_u=1))):
        w = make()
        lat = w["lattice"]
        comm = TorchComm(periodic_x=periodic_x, device=f"cuda:{local}")
        s = W.make_solver(w, comm=comm, seed=77, device=local, options=opts)
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
            single = W.make_solver(w, seed=77, device=local, options=opts)
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
    flag = torch.tensor([1 if ok else 0], device=f"cuda:{local}")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    if rank == 0:
        print("MGPU_CHECK_OK" if ok else "MGPU_CHECK_FAILED", flush=True)
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
