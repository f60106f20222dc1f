"""Development helper: time the two stage launches for a list of option sets (not the contract bench; see bench.py).

    python scripts/quick_bench.py [--workload c3|c2|c4] [--dims 256x256x256] [--T 0,100] [--steps 20] [--trace] ['{"recover_u": 0}' '{"chunks": 9}' ...]

Every option set runs in its own process (a CUDA error is sticky for the process that hit it).  With --trace the per-CTA
busy times of the last stage-A / stage-B launch are summarised (min / mean / max, items per CTA)."""
import json
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

PEAK = 6545.3e9
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] * 1e9
except Exception:
    pass


def trace_summary(ctx, tag, dump=None):
    t = ctx.last_stage_trace()
    if len(t) == 0:
        return
    if dump:
        np.save(dump, t)
    busy = (t[:, 2] - t[:, 1]).astype(np.float64) * 1e-3   # us
    span = float(t[:, 2].max() - t[:, 1].min()) * 1e-3
    items = t[:, 3].astype(np.int64)
    sm = t[:, 0].astype(np.int64)
    per_sm = {}
    for k in range(len(t)):
        per_sm.setdefault(int(sm[k]), []).append(float(t[k, 2]))
    last_end = np.array([max(v) for v in per_sm.values()]) - float(t[:, 1].min())
    print(f"    trace {tag}: {len(t)} CTAs on {len(per_sm)} SMs, span {span:.1f} us; CTA busy min/mean/max {busy.min():.1f}/{busy.mean():.1f}/{busy.max():.1f} us "
          f"(mean/span {busy.mean() / span:.3f}); SM end-time min/mean {last_end.min() * 1e-3:.1f}/{last_end.mean() * 1e-3:.1f} us; items per CTA min/max {items.min()}/{items.max()}",
          flush=True)


def run(dims, options, T, steps, label, trace):
    from jams_b200 import workloads as W
    kind = os.environ.get("JB_QB_WORKLOAD", "c3")     # --workload c2 | c4: cubic bcc lattices of dims[0]^3
    w = W.c3_sc(dims=dims, temperature=T) if kind == "c3" else (W.c2_bcc_fe(dims[0], temperature=T) if kind == "c2" else W.c4_bcc_long_range(dims[0], temperature=T))
    try:
        s = W.make_solver(w, options=dict(options, time_kernels=1, trace=1 if trace else 0), random_spins_seed=1)
        s.run(3); s.ctx.synchronize(); s.ctx.last_step_kernel_ms()
        t0 = time.perf_counter(); s.run(steps); s.ctx.synchronize(); wall = time.perf_counter() - t0
        ms = s.ctx.last_step_kernel_ms()
    except Exception as e:  # noqa: BLE001
        print(f"{label:60s} FAILED: {e}", flush=True)
        return
    N = w["lattice"].num_spins
    rate = N * steps / (ms.sum() * 1e-3)
    print(f"{label:60s} T={T:5.0f}: A {ms[0]/steps:.4f} ms  B {ms[1]/steps:.4f} ms  wall/step {wall/steps*1e3:.4f} ms "
          f"-> {rate/1e9:6.2f} G upd/s = {rate*144/PEAK*100:5.1f}% of HBM roofline (144 B model)", flush=True)
    if trace:
        dump = None
        if os.environ.get("JB_TRACE_DUMP"):
            dump = os.path.join(os.environ["JB_TRACE_DUMP"], "trace_T%d_%s.npy" % (int(T), "".join(ch if ch.isalnum() else "_" for ch in label)[:60]))
        trace_summary(s.ctx, "last launch (stage B)", dump)
    s.ctx.close()


if __name__ == "__main__":
    args = sys.argv[1:]
    if args and args[0] == "--one":
        dims = tuple(int(v) for v in args[1].split("x"))
        run(dims, json.loads(args[4]), float(args[2]), int(args[3]), args[4], args[5] == "1")
        sys.exit(0)
    dims, temps, steps, trace, sets = "256x256x256", [0.0, 100.0], 20, False, []
    i = 0
    while i < len(args):
        if args[i] == "--dims": dims = args[i + 1]; i += 2
        elif args[i] == "--T": temps = [float(v) for v in args[i + 1].split(",")]; i += 2
        elif args[i] == "--steps": steps = int(args[i + 1]); i += 2
        elif args[i] == "--trace": trace = True; i += 1
        elif args[i] == "--workload": os.environ["JB_QB_WORKLOAD"] = args[i + 1]; i += 2
        else: sets.append(args[i]); i += 1
    if not sets:
        sets = ["{}"]
    for T in temps:
        for o in sets:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", dims, str(T), str(steps), o, "1" if trace else "0"],
                               capture_output=True, text=True, timeout=300)
            for line in (r.stderr or "").splitlines():
                if line.startswith("jams_b200:"):
                    print("    " + line[:400], flush=True)
            out = (r.stdout or "").strip()
            print(out if out else f"{o:60s} CRASHED rc={r.returncode}: {(r.stderr or '').strip()[-400:]}", flush=True)
