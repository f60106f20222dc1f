"""Where does the multi-GPU step lose time?  Every rank times the two stage kernels of (a) its own single-slab 256^3 lattice and
(b) its slab of the (256 N) x 256 x 256 lattice, then dumps the per-item CTA trace of the last stage launches of (b):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/mgpu_trace.py [json options]
Face items (the two chunks that touch the slab's x faces) against interior items: start time, duration per plane."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from jams_b200 import capi, workloads as W
from jams_b200.distributed import TorchComm


def item_table(trace, n_cols, plan, gx, nx):
    rows = []
    for cta in trace:
        n = int(cta[3])
        t_end = float(cta[2] - cta[1])
        starts = [(int(v >> 40), float(v & 0xffffffffff)) for v in cta[4:4 + min(n, 28)]]
        for k, (item, t0) in enumerate(starts):
            t1 = starts[k + 1][1] if k + 1 < len(starts) else (t_end if n <= 28 else np.nan)
            chunk = item // n_cols
            x0, xc = plan[chunk]
            face = (x0 < gx) or (x0 + xc > nx - gx)
            rows.append((item, chunk, xc, face, t0 * 1e-3, (t1 - t0) * 1e-3, float(cta[1])))
    return rows


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    extra = json.loads(sys.argv[1]) if len(sys.argv) > 1 else {}
    T, steps = 100.0, 60
    res = {}
    for name, dims, comm in (("single", (256, 256, 256), None), ("slab", (256 * world, 256, 256), TorchComm(periodic_x=True, device=f"cuda:{local}"))):
        w = W.c3_sc(dims=dims, temperature=T)
        s = W.make_solver(w, comm=comm, seed=3, device=local, options=dict(extra, time_kernels=1, trace=1), random_spins_seed=1)
        s.run(10); s.ctx.synchronize(); s.ctx.last_step_kernel_ms()
        if comm: comm.barrier(s.ctx)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st = torch.cuda.ExternalStream(s.ctx.stream())
        ev0.record(st); s.run(steps); ev1.record(st); s.ctx.synchronize()
        ms = s.ctx.last_step_kernel_ms() / steps
        res[name] = (ms, ev0.elapsed_time(ev1) / steps)
        if name == "slab":
            tr = s.ctx.last_stage_trace()
            n_cols = (256 // extra.get("tile_y", 4)) * (256 // extra.get("tile_z", 64))   # the library's default tile here is 4 x 64
            try:
                plan = capi.plan_work_items(256, 1, n_cols, len(tr))
                rows = item_table(tr, n_cols, plan, 1, 256)
            except Exception as e:  # noqa: BLE001  (a plan the host-side helper does not reproduce: options that change it)
                print(f"[rank {rank}] per-item table not available: {e}", flush=True)
                rows = None
        if name == "slab" and rows:
            face = [r for r in rows if r[3]]; inner = [r for r in rows if not r[3] and r[2] >= 8]
            per_plane = lambda rs: np.nanmean([r[5] / (r[2] + 2) for r in rs])
            span = float(tr[:, 2].max() - tr[:, 1].min()) * 1e-3
            busy = (tr[:, 2] - tr[:, 1]).astype(float) * 1e-3
            print(f"[rank {rank}] last launch (stage B): span {span:.1f} us, CTA busy mean {busy.mean():.1f} max {busy.max():.1f}; launch skew of CTA starts {(tr[:, 1].max() - tr[:, 1].min()) * 1e-3:.1f} us; "
                  f"face items {len(face)}: start {np.mean([r[4] for r in face]):.1f} us, duration {np.nanmean([r[5] for r in face]):.1f} us, {per_plane(face):.3f} us per plane; "
                  f"interior items {len(inner)}: {per_plane(inner):.3f} us per plane", flush=True)
        if comm: comm.barrier(s.ctx)
        s.ctx.close()
    a, b = res["single"], res["slab"]
    print(f"[rank {rank}] single slab 256^3: A {a[0][0]:.4f} B {a[0][1]:.4f} ms, step {a[1]:.4f} ms | slab of {world}: A {b[0][0]:.4f} B {b[0][1]:.4f} ms, step {b[1]:.4f} ms "
          f"-> efficiency {a[1] / b[1]:.3f}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
