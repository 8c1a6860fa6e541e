#!/usr/bin/env python
"""bench.py — CP-ALS sweeps/s and MTTKRP HBM GB/s on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload c2|c5|c1|small] [--no-e2e]

A "step" is one CP-ALS sweep (all modes updated + reconstruction error) of the workload:
  c2 (default): parafac rank 32 on random 1024^3 fp32  (BASELINE configs[1])
  c5          : parafac rank 64 on random 2048^3 fp32  (BASELINE configs[4])
At N > 1 the SAME tensor is sharded along mode 0 over the N ranks (strong scaling); the
per-mode exchanges are NCCL all-reduces of the small MTTKRP / Gram partials.

One JSON line is printed by rank 0 (see DESIGN.md §Measurement for every key).
`--impl reference` times the reference's own CPU implementation (unmodified TensorLy on
its numpy backend when importable from baseline/_ref, else the oracle port) on a bounded
sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    "c2": dict(shape=(1024, 1024, 1024), rank=32, dtype="float32",
               name="C2: parafac CP-ALS rank 32 on random 1024x1024x1024 fp32"),
    "c5": dict(shape=(2048, 2048, 2048), rank=64, dtype="float32",
               name="C5: parafac CP-ALS rank 64 on random 2048x2048x2048 fp32 (sharded along mode 0 when N>1)"),
    "c1": dict(shape=(100, 100, 100), rank=10, dtype="float64",
               name="C1: parafac CP-ALS rank 10 on random 100x100x100 float64"),
    "small": dict(shape=(256, 256, 256), rank=32, dtype="float32", name="small: parafac rank 32 on 256^3 fp32"),
}
METRIC = "CP-ALS sweeps/s (MTTKRP HBM GB/s in roofline)"
UNIT = "sweeps/s"


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed regions run."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index), "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            parts = [p.strip() for p in r.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); smax.append(float(parts[1])); power.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        # "under load": samples in the upper half of the observed power range
        if sm:
            thr = (max(power) + min(power)) / 2 if power else 0
            loaded = [s for s, p in zip(sm, power) if p >= thr] or sm
            return {"sm_mhz": statistics.median(loaded), "sm_max_mhz": max(smax), "reasons": sorted(reasons),
                    "samples": len(sm), "power_w_max": max(power) if power else None}
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"], "samples": 0}


# ----------------------------------------------------------------------------------------
def cpu_reference_sweeps(shape, rank, dtype, steps, warmup):
    """Time the reference's CPU implementation (numpy backend, `core` tenalg) for
    `steps` sweeps after `warmup`, fixed init, tol=0, errors evaluated.  Returns
    (seconds_per_sweep, kind, threads)."""
    import numpy as np
    from oracle import oracle as O
    x = O.random_tensor(shape, 0, np.dtype(dtype))
    w, fs = O.random_cp_factors(shape, rank, 1, np.dtype(dtype))
    try:
        from threadpoolctl import threadpool_info
        threads = max([p.get("num_threads", 1) for p in threadpool_info()] or [os.cpu_count() or 1])
    except Exception:
        threads = os.cpu_count() or 1
    kind = "port"
    run = None
    try:
        from tensorly_b200.backend import import_tensorly
        tl = import_tensorly()
        tl.set_backend("numpy")
        tl.tenalg.set_backend("core")
        from tensorly.cp_tensor import CPTensor
        from tensorly.decomposition import parafac

        def run(n):
            init = CPTensor((w.copy(), [f.copy() for f in fs]))
            t0 = time.perf_counter()
            parafac(x, rank, n_iter_max=n, init=init, tol=0, return_errors=True)
            return time.perf_counter() - t0
        kind = "reference"
    except Exception:
        def run(n):
            t0 = time.perf_counter()
            O.parafac(x, (w, fs), n_iter_max=n)
            return time.perf_counter() - t0
    # sweeps/s = delta(iters)/delta(time) between two runs removes init + tl.norm
    t_a = run(warmup)
    t_b = run(warmup + steps)
    per = max((t_b - t_a) / steps, 1e-9)
    return per, kind, threads


def run_reference(args):
    rank_env = int(os.environ.get("RANK", "0"))
    if rank_env != 0:
        return 0
    wl = WORKLOADS[args.workload]
    shape, rank, dtype = wl["shape"], wl["rank"], wl["dtype"]
    # bounded sample: the same problem at a sub-shape that a CPU sweeps in ~a second
    sample = tuple(min(s, 512 if args.workload in ("c2", "c5") else s) for s in shape)
    frac = 1.0
    for a, b in zip(sample, shape):
        frac *= a / b
    steps = max(1, min(args.steps, 8))
    warm = max(1, min(args.warmup, 2))
    per, kind, threads = cpu_reference_sweeps(sample, rank, dtype, steps, warm)
    value = frac / per     # MTTKRP cost is linear in the element count
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": per / frac * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32" if dtype == "float32" else "f64", "data": "synthetic",
        "config": {"workload": wl["name"], "rank": rank, "shape": list(shape)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": f"{steps} sweeps of the same CP-ALS at sub-shape {sample} (rank {rank}, {dtype}), "
                                   f"scaled by the element ratio {frac:.6g}"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import tensorly_b200 as tb
    from tensorly_b200.cp_als import CPALS, _Comm

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank_id = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # the sharded sweep (MTTKRP + NCCL all-reduce + solve, 3 modes) is replayed from one CUDA graph;
        # the graph is destroyed before the process group at the end (the reverse order hangs in teardown)
        os.environ.setdefault("TLB200_DIST_GRAPH", "1")
        dist.init_process_group("nccl", device_id=device)
    wl = WORKLOADS[args.workload]
    shape, R = wl["shape"], wl["rank"]
    dtype = torch.float32 if wl["dtype"] == "float32" else torch.float64
    esize = 4 if dtype == torch.float32 else 8
    lo, hi = tb.shard_bounds(shape[0], world, rank_id)
    local_shape = (hi - lo,) + tuple(shape[1:])

    # synthetic data: uniform[0,1) like tl.random.random_tensor, generated on device
    gen = torch.Generator(device=device).manual_seed(1234 + rank_id)
    x = torch.rand(local_shape, generator=gen, dtype=dtype, device=device)
    fgen = torch.Generator(device=device).manual_seed(1)
    factors = [torch.rand((s, R), generator=fgen, dtype=dtype, device=device) for s in shape]
    factors[0] = factors[0][lo:hi].contiguous()
    weights = torch.ones(R, dtype=dtype, device=device)
    comm = _Comm(None)
    state = CPALS(x, weights, factors, comm=comm, shard_mode=0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank_id == 0:
        sampler.start()

    # ---- roofline leg: the MTTKRP call per mode, timed with CUDA events on its stream -----
    elems_local = 1
    for s in local_shape:
        elems_local *= s
    alg_bytes = esize * (elems_local + R * sum(local_shape))
    mttkrp_ms = []
    path_used = None
    for mode in range(len(shape)):
        for _ in range(3):
            tb.unfolding_dot_khatri_rao(x, (weights, state.factors), mode)
        path_used = tb.last_kernel_path()
        reps = max(3, min(20, args.steps))
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
        torch.cuda.synchronize()
        ev[0].record()
        for i in range(reps):
            tb.unfolding_dot_khatri_rao(x, (weights, state.factors), mode)
            ev[i + 1].record()
        torch.cuda.synchronize()
        mttkrp_ms.append(statistics.mean(ev[i].elapsed_time(ev[i + 1]) for i in range(reps)))
    avg_ms = statistics.mean(mttkrp_ms)
    achieved = alg_bytes / (avg_ms * 1e-3) / 1e9
    peak, peak_src = measured_peak_hbm()

    # ---- launches per sweep (host-side count of one eager sweep) -------------------------
    c0 = tb.launch_count()
    state.sweep_eager(True)
    launches_per_sweep = tb.launch_count() - c0

    # ---- timed region: K sweeps, inputs resident in HBM ----------------------------------
    for _ in range(max(3, args.warmup)):
        state.sweep(True)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        state.sweep(True)
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    value = args.steps / (total_ms * 1e-3)
    rel_err = float(state.err[0].item())

    # ---- the same sweeps with one full MTTKRP per mode (N tensor passes, the reference's call pattern) ----
    three_pass = None
    ttm_pass = None
    if state.dimtree:
        st3 = CPALS(x, weights, factors, comm=comm, shard_mode=0, dimtree=False)
        n3 = max(3, min(args.steps, 30))
        for _ in range(3):
            st3.sweep(True)
        barrier()
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record()
        for _ in range(n3):
            st3.sweep(True)
        b1.record()
        barrier()
        ms3 = torch.tensor([b0.elapsed_time(b1)], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(ms3, op=dist.ReduceOp.MAX)
        three_pass = {"value": n3 / (float(ms3.item()) * 1e-3), "unit": UNIT, "ms_per_step": float(ms3.item()) / n3,
                      "steps": n3, "final_rel_error": float(st3.err[0].item()),
                      "what": "same sweeps with a full MTTKRP per mode (no dimension-tree reuse): the call pattern "
                              "SURVEY 8(d)'s 12.89 GB/sweep model describes"}
        st3._graph = None
        del st3
        # the TTM pass that feeds the dimension tree (same tcgen05 engine): bytes = tensor read + T written
        last = len(shape) - 1
        for _ in range(3):
            tb.mode_dot(x, state.factors[last], last, transpose=True)
        c0_, c1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        c0_.record()
        for _ in range(5):
            tb.mode_dot(x, state.factors[last], last, transpose=True)
        c1_.record()
        torch.cuda.synchronize()
        t_ms = c0_.elapsed_time(c1_) / 5
        t_bytes = esize * (elems_local + elems_local // local_shape[last] * R + R * local_shape[last])
        ttm_pass = {"ms_per_launch": t_ms, "algorithmic_bytes_per_launch": t_bytes, "achieved": t_bytes / (t_ms * 1e-3) / 1e9,
                    "unit": "GB/s", "kernel": f"mode_dot(X, F_last^T, last) ({tb.last_kernel_path()})"}

    # ---- e2e: same sweep through the public API with HOST buffers ------------------------
    e2e = None
    if not args.no_e2e:
        x_host = torch.empty(local_shape, dtype=dtype, pin_memory=True)
        x_host.copy_(x)
        f_host = [torch.empty(f.shape, dtype=dtype, pin_memory=True).copy_(f) for f in state.factors]
        out_host = [torch.empty(f.shape, dtype=dtype, pin_memory=True) for f in state.factors]
        err_host = torch.empty(3, dtype=dtype, pin_memory=True)
        x_dev = torch.empty_like(x)
        e2e_steps = max(2, min(args.steps, 6))

        def e2e_step():
            x_dev.copy_(x_host, non_blocking=True)                        # H2D: the tensor
            fs = [h.to(device, non_blocking=True) for h in f_host]        # H2D: current factors
            st = CPALS(x_dev, weights, fs, comm=comm, shard_mode=0)       # ||X||^2 + Grams
            st.sweep_eager(True)
            for o, f in zip(out_host, st.factors):
                o.copy_(f, non_blocking=True)                             # D2H: updated factors
            err_host.copy_(st.err, non_blocking=True)                     # D2H: the error
            torch.cuda.current_stream().synchronize()

        for _ in range(2):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(e2e_steps):
            e2e_step()
        a1.record()
        barrier()
        t = torch.tensor([max(a0.elapsed_time(a1) * 1e-3, time.perf_counter() - t0)], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        fbytes = esize * R * sum(local_shape)
        e2e = {"value": e2e_steps / float(t.item()), "unit": UNIT,
               "h2d_bytes_per_step": int(esize * elems_local + fbytes) * world,
               "d2h_bytes_per_step": int(fbytes + 3 * esize) * world, "steps": e2e_steps,
               "note": "per step: pinned-host tensor slab + factors -> device, ||X||^2 + Grams + one full ALS sweep "
                       "through tensorly_b200.CPALS, factors + error -> host"}
        del x_host, x_dev

    clocks = sampler.stop() if rank_id == 0 else None

    # ---- optional: the UNMODIFIED reference driver on the b200 tenalg backend -------------
    ref_driver = None
    if rank_id == 0 and world == 1 and not args.no_refdriver:
        try:
            tl = tb.import_tensorly()
            tl.set_backend("pytorch")
            tb.use()
            from tensorly.cp_tensor import CPTensor
            from tensorly.decomposition import parafac

            def timed(n):
                init = CPTensor((weights.clone(), [f.clone() for f in factors]))
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                parafac(x, R, n_iter_max=n, init=init, tol=0, return_errors=True)
                torch.cuda.synchronize()
                return time.perf_counter() - t0
            timed(2)
            ta, tb_ = timed(2), timed(12)
            ref_driver = {"value": 10.0 / max(tb_ - ta, 1e-9), "unit": UNIT,
                          "what": "tensorly.decomposition.parafac (unmodified) on tl.tenalg backend 'b200'"}
            tb.set_dimension_tree(True)
            try:
                timed(2)
                ta, tb_ = timed(2), timed(12)
                ref_driver["value_dimension_tree"] = 10.0 / max(tb_ - ta, 1e-9)
                ref_driver["what_dimension_tree"] = "same, with tensorly_b200.use(dimension_tree=True)"
            finally:
                tb.set_dimension_tree(False)
        except Exception as exc:  # tensorly not importable on this box
            ref_driver = {"unavailable": str(exc)[:200]}

    # ---- CPU baseline on the host cores (rank 0, N=1 only), bounded sample ---------------
    cpu = None
    if rank_id == 0 and world == 1 and not args.no_cpu:
        sample = tuple(min(s, 512 if args.workload in ("c2", "c5") else s) for s in shape)
        frac = 1.0
        for a, b in zip(sample, shape):
            frac *= a / b
        per, kind, threads = cpu_reference_sweeps(sample, R, wl["dtype"], 3, 1)
        cpu = {"value": frac / per, "unit": UNIT, "cores": threads, "kind": kind,
               "sample": f"3 sweeps of the same CP-ALS at sub-shape {sample} (rank {R}, {wl['dtype']}), numpy backend + "
                         f"core tenalg, scaled by the element ratio {frac:.6g}"}

    if rank_id == 0:
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if world == 1 and os.path.exists(tp):      # the ncu capture is of the 1-GPU launch of this workload
            try:
                traffic = json.load(open(tp)).get(args.workload, {}).get(path_used)
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32" if dtype == torch.float32 else "f64", "data": "synthetic",
            "config": {"workload": wl["name"], "shape": list(shape), "rank": R, "sharding": f"mode-0 slabs over {world} GPU(s)",
                       "l2": "inputs larger than L2 (tensor slab %.2f GB per GPU >> 126 MB)" % (esize * elems_local / 1e9),
                       "kernel_path": path_used,
                       "sweep": ("dimension-tree ALS sweep: T = X x_last F_last^T (one tensor pass) -> MTTKRP of every earlier "
                                 "mode from T, full MTTKRP for the last mode (second tensor pass); identical factor updates, "
                                 "parity-tested against the N-pass sweep" if state.dimtree else
                                 "one full MTTKRP per mode") +
                                (" + Gram-Hadamard LU solve + Gram per mode + error, one CUDA graph" if world == 1 else
                                 " + all_reduce + solve + Gram per mode + error, one CUDA graph")},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": f"MTTKRP ({path_used})",
                         "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": avg_ms,
                         "per_mode_gbs": [alg_bytes / (m * 1e-3) / 1e9 for m in mttkrp_ms],
                         "frac_of_nominal_8TBs": achieved / 8000.0,
                         "note": ("frac > 1: `peak` is the measured COPY bandwidth (read + write); a read-only stream like "
                                  "this one can exceed it — see frac_of_nominal_8TBs") if achieved > peak else None},
            "three_pass": three_pass,
            "ttm_pass": ttm_pass,
            "cpu_baseline": cpu,
            "e2e": e2e,
            "gpu_launches": int(launches_per_sweep * args.steps),
            "launches_per_sweep": int(launches_per_sweep),
            "clocks": clocks,
            "final_rel_error": rel_err,
            "reference_driver_on_b200": ref_driver,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        # teardown must never turn a finished measurement into a hang: drop the CUDA graphs (they hold NCCL
        # kernels) before the communicator, and bail out hard if the collective teardown stalls anyway
        import gc
        def _bail():
            time.sleep(30)
            os._exit(0)
        threading.Thread(target=_bail, daemon=True).start()
        state._graph = None
        del state
        gc.collect()
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-refdriver", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
