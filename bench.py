#!/usr/bin/env python
"""bench.py — CP-ALS sweeps/s and MTTKRP HBM GB/s on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload c5|c2|c1|small] [--no-e2e] [--no-c2] [--no-cpu] ...

A "step" is one CP-ALS sweep (all modes updated + reconstruction error) of the workload.
The headline workload is the same at EVERY N so that the driver can compute scaling from
the per-N values:

  c5 (default): parafac rank 64 on random 2048^3 fp32  (BASELINE configs[4], the config the
                north star's 8-GPU target is quoted on; 34.4 GB, fits one B200).  At N > 1 the
                SAME tensor (generated block by block from fixed seeds) is sharded along mode 0
                over the N ranks: strong scaling.
  c2 block    : parafac rank 32 on random 1024^3 fp32  (BASELINE configs[1]) runs in the same
                process as a secondary block `c2` with its own value, roofline and — at N = 1 —
                its own CPU baseline measured on the REAL 1024^3 shape.

One JSON line is printed by rank 0 (DESIGN.md "Measurement" explains every key).

`--impl reference` times the UNMODIFIED reference (TensorLy from baseline/_ref, numpy backend,
`core` tenalg; the oracle port only if TensorLy is not importable) on the host cores, with every
host thread, on the same config: real sweeps of the real shape, each sweep time-stamped through the
reference's own `verbose` prints.  K and W are honoured as far as a wall-clock budget allows
(`--ref-budget-s`, default 180 s): the line's `steps` / `warmup` are the sweeps that were actually
timed, so ms_per_step x steps is real time.  Only when the host cannot hold the reference's
temporaries (C5 needs ~3.2x the 34 GB tensor) a mode-0 slab is timed instead and labelled
`extrapolated`.
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
    # one rank's share of C5 at N = 8 as a single-GPU problem (profiling aid: same kernels, no exchange)
    "c5slab": dict(shape=(256, 2048, 2048), rank=64, dtype="float32", name="C5 slab: 256x2048x2048 rank 64 (1/8 of C5)"),
    "c2slab": dict(shape=(128, 1024, 1024), rank=32, dtype="float32", name="C2 slab: 128x1024x1024 rank 32 (1/8 of C2)"),
}
METRIC = "CP-ALS sweeps/s (MTTKRP HBM GB/s in roofline)"
UNIT = "sweeps/s"
PARITY_SHAPE, PARITY_RANK, PARITY_SWEEPS = (384, 320, 256), 32, 6


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def host_mem_available_bytes() -> int:
    """MemAvailable, capped by the cgroup limit when there is one."""
    avail = None
    try:
        with open("/proc/meminfo") as f:
            for line in f:
                if line.startswith("MemAvailable:"):
                    avail = int(line.split()[1]) * 1024
                    break
    except Exception:
        pass
    for p, q in (("/sys/fs/cgroup/memory.max", "/sys/fs/cgroup/memory.current"),
                 ("/sys/fs/cgroup/memory/memory.limit_in_bytes", "/sys/fs/cgroup/memory/memory.usage_in_bytes")):
        try:
            lim = open(p).read().strip()
            if lim != "max" and int(lim) < (1 << 60):
                room = int(lim) - int(open(q).read().strip())
                avail = room if avail is None else min(avail, room)
        except Exception:
            pass
    return avail if avail is not None else 0


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
         "clocks_event_reasons.sw_power_cap,clocks.mem,temperature.gpu")

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

    def mark(self):
        return len(self.rows)

    def summary(self, start=0, end=None):
        sm, smax, reasons, power, mem, temp = [], [], set(), [], [], []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows[start:end]:
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
            try:
                mem.append(float(parts[7])); temp.append(float(parts[8]))
            except (IndexError, ValueError):
                pass
        if sm:
            # "under load": samples in the upper half of the observed power range
            thr = (max(power) + min(power)) / 2 if power else 0
            loaded = [s for s, p in zip(sm, power) if p >= thr] or sm
            return {"sm_mhz": statistics.median(loaded), "sm_max_mhz": max(smax), "reasons": sorted(reasons),
                    "samples": len(sm), "power_w_max": max(power) if power else None,
                    "mem_mhz": statistics.median(mem) if mem else None, "temp_c_max": max(temp) if temp else None}
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"], "samples": 0}

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        return self.summary()


# ---------------------------------------------------------------------------------------- CPU reference
def _full_threads():
    """Context manager forcing BLAS/OpenMP pools to every host thread (torchrun exports OMP_NUM_THREADS=1)."""
    n = host_threads()
    try:
        from threadpoolctl import threadpool_limits
        return threadpool_limits(limits=n), n
    except Exception:
        import contextlib
        return contextlib.nullcontext(), n


def _blas_threads(default):
    try:
        from threadpoolctl import threadpool_info
        return max([p.get("num_threads", 1) for p in threadpool_info()] or [default])
    except Exception:
        return default


def host_tensor(shape, dtype, seed=0):
    """uniform[0,1) like tl.random.random_tensor; generated with torch's CPU generator in mode-0 blocks (numpy's
    RandomState would need an fp64 copy of the whole tensor: 69 GB at C5)."""
    import numpy as np
    import torch
    tdt = torch.float32 if dtype == "float32" else torch.float64
    x = torch.empty(shape, dtype=tdt)
    g = torch.Generator().manual_seed(seed)
    step = max(1, (1 << 28) // max(1, int(np.prod(shape[1:]))))
    for lo in range(0, shape[0], step):
        x[lo:lo + step].uniform_(0, 1, generator=g)
    return x.numpy()


class _StampingStdout:
    """Time-stamps the reference's own `verbose` prints ("Starting iteration k", _cp.py:404-405): per-sweep wall
    times of the unmodified driver without its `callback` hook (which would add a full cp_to_tensor
    reconstruction — two more tensor-sized temporaries — at start-up, _cp.py:383-393)."""

    def __init__(self):
        self.stamps = []

    def write(self, text):
        if text.startswith("Starting iteration"):
            self.stamps.append(time.perf_counter())
        return len(text)

    def flush(self):
        pass


def _reference_parafac_timed(x, rank, w, fs, n_iter):
    """One call of the unmodified reference parafac for n_iter sweeps; returns the start time of every sweep plus
    the end time, and which implementation ran."""
    import contextlib
    try:
        from tensorly_b200.backend import import_tensorly
        tl = import_tensorly()
    except ImportError:
        tl = None
    if tl is not None:
        tl.set_backend("numpy")
        tl.tenalg.set_backend("core")
        from tensorly.cp_tensor import CPTensor
        from tensorly.decomposition import parafac
        init = CPTensor((w.copy(), [f.copy() for f in fs]))
        out = _StampingStdout()
        with contextlib.redirect_stdout(out):
            parafac(x, rank, n_iter_max=n_iter, init=init, tol=0, return_errors=True, verbose=2)
        return out.stamps + [time.perf_counter()], "reference"
    # TensorLy not importable: the oracle port, one sweep per call from the running factors
    from oracle import oracle as O
    cur = (w.copy(), [f.copy() for f in fs])
    stamps = []
    for _ in range(n_iter):
        stamps.append(time.perf_counter())
        (wn, fn), _ = O.parafac(x, cur, n_iter_max=1)
        cur = (wn, fn)
    return stamps + [time.perf_counter()], "port"


def cpu_reference_sweeps(shape, rank, dtype, steps, warmup, budget_s, min_steps=2):
    """Sweeps of the reference's CPU CP-ALS (unmodified parafac, numpy backend, `core` tenalg, fixed init, tol=0,
    errors evaluated) at `shape`, every host thread.  `warmup` untimed + `steps` timed sweeps when that fits the
    wall-clock budget; otherwise fewer (at least 1 + `min_steps`), decided up front from a calibration sweep on a
    thin mode-0 slab.  Returns (seconds_per_sweep, steps_timed, warmup_done, kind, threads)."""
    import numpy as np
    from oracle import oracle as O
    ctx, nthreads = _full_threads()
    with ctx:
        t_begin = time.perf_counter()
        x = host_tensor(shape, dtype, 0)
        w, fs = O.random_cp_factors(shape, rank, 1, np.dtype(dtype))
        threads = _blas_threads(nthreads)
        inner = 1
        for s in shape[1:]:
            inner *= s
        rows = max(1, min(shape[0], (1 << 27) // max(inner, 1)))
        if rows < shape[0]:
            st, _ = _reference_parafac_timed(np.ascontiguousarray(x[:rows]), rank, w, [fs[0][:rows]] + list(fs[1:]), 2)
            est = (st[-1] - st[1]) * shape[0] / rows          # second sweep (the first one warms the BLAS pools)
        else:
            est = 0.0
        left = budget_s - (time.perf_counter() - t_begin)
        warm, timed = warmup, steps
        if est * (warm + timed) > left:
            total = max(1 + min_steps, int(left / max(est, 1e-9)))
            warm, timed = 1, max(min_steps, min(steps, total - 1))
        stamps, kind = _reference_parafac_timed(x, rank, w, fs, warm + timed)
    per = (stamps[-1] - stamps[warm]) / timed
    return per, timed, warm, kind, threads


def reference_plan(shape, rank, dtype):
    """The shape the CPU arm can really run: the workload's own shape when the reference's temporaries fit in
    host memory (tensor + permuted unfolding copy + the |x|, x^2 temporaries of tl.norm ~ 3.2x the tensor), else
    the largest mode-0 slab that does."""
    esize = 4 if dtype == "float32" else 8
    elems = 1
    for s in shape:
        elems *= s
    need = 3.2 * esize * elems + (2 << 30)
    avail = host_mem_available_bytes()
    if avail <= 0 or need <= avail:
        return tuple(shape), 1.0, avail
    rows = shape[0]
    while rows > 1 and 3.2 * esize * elems * rows / shape[0] + (2 << 30) > avail:
        rows //= 2
    return (rows,) + tuple(shape[1:]), rows / shape[0], avail


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    wl = WORKLOADS[args.workload]
    shape, rank, dtype = wl["shape"], wl["rank"], wl["dtype"]
    sample, frac, avail = reference_plan(shape, rank, dtype)
    per, steps, warm, kind, threads = cpu_reference_sweeps(sample, rank, dtype, args.steps, args.warmup, args.ref_budget_s)
    extrapolated = frac != 1.0
    value = frac / per          # slab sample: MTTKRP cost is linear in the slab's rows
    what = (f"{steps} timed sweeps (after {warm} warm-up) of unmodified tensorly.decomposition.parafac "
            f"(numpy backend, core tenalg) at shape {tuple(sample)} rank {rank} {dtype}, {threads} threads, "
            f"per-sweep times from the time-stamped `verbose` prints of the driver")
    if extrapolated:
        what += (f"; EXTRAPOLATED: host memory ({avail / 2**30:.0f} GiB available) cannot hold the reference's "
                 f"temporaries for {tuple(shape)}, so a mode-0 slab was timed and scaled by {frac:.4g}")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "steps_requested": args.steps, "warmup_requested": args.warmup,
        "ms_per_step": per * 1e3 / frac, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32" if dtype == "float32" else "f64", "data": "synthetic",
        "config": {"workload": wl["name"], "shape": list(shape), "rank": rank},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": what,
                         "extrapolated": extrapolated, "measured_ms_per_sweep_at_sample": per * 1e3,
                         "sample_shape": list(sample)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------- synthetic data
def device_slab(shape, lo, hi, dtype, device, seed=0, block=64):
    """Rows lo:hi (mode 0) of THE synthetic tensor: uniform[0,1) like tl.random.random_tensor, produced in mode-0
    blocks of `block` rows, each from its own seeded device generator — the same tensor whatever the number of
    ranks and wherever the slab boundaries fall."""
    import torch
    x = torch.empty((hi - lo,) + tuple(shape[1:]), dtype=dtype, device=device)
    g = torch.Generator(device=device)
    for b in range(lo // block, (hi + block - 1) // block):
        b0, b1 = b * block, min((b + 1) * block, shape[0])
        g.manual_seed(seed * 1000003 + b)
        blk = torch.rand((b1 - b0,) + tuple(shape[1:]), generator=g, dtype=dtype, device=device)
        s0, s1 = max(b0, lo), min(b1, hi)
        x[s0 - lo:s1 - lo] = blk[s0 - b0:s1 - b0]
        del blk
    return x


def initial_factors(shape, rank, dtype, device, seed=1):
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    return [torch.rand((s, rank), generator=g, dtype=dtype, device=device) for s in shape]


# ---------------------------------------------------------------------------------------- our arm
class Env:
    pass


def timed_sweeps(env, state, steps, warmup):
    """`steps` sweeps between barrier+synchronize brackets, CUDA events on the launching stream, max over ranks."""
    import torch
    for _ in range(warmup):
        state.sweep(True)
    env.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        state.sweep(True)
    e1.record()
    env.barrier()
    return env.max_over_ranks(e0.elapsed_time(e1))


def bench_workload(env, key, args, headline):
    """All device-resident measurements of one workload: MTTKRP roofline leg, K timed sweeps, the N-pass variant,
    the TTM pass that feeds the dimension tree, launches per sweep."""
    import torch
    import tensorly_b200 as tb
    from tensorly_b200.cp_als import CPALS
    wl = WORKLOADS[key]
    shape, R = wl["shape"], wl["rank"]
    dtype = torch.float32 if wl["dtype"] == "float32" else torch.float64
    esize = 4 if dtype == torch.float32 else 8
    lo, hi = tb.shard_bounds(shape[0], env.world, env.rank)
    local_shape = (hi - lo,) + tuple(shape[1:])
    x = device_slab(shape, lo, hi, dtype, env.device, seed=0)
    factors = initial_factors(shape, R, dtype, env.device)
    factors[0] = factors[0][lo:hi].contiguous()
    weights = torch.ones(R, dtype=dtype, device=env.device)
    state = CPALS(x, weights, factors, comm=env.comm, shard_mode=0)
    out = {"x": x, "weights": weights, "factors": factors, "state": state, "local_shape": local_shape, "esize": esize}

    # ---- roofline leg: the MTTKRP call per mode, timed with CUDA events on its stream -----
    elems_local = 1
    for s in local_shape:
        elems_local *= s
    alg_bytes = esize * (elems_local + R * sum(local_shape))
    mttkrp_ms, path_used = [], None
    reps = max(3, min(20, args.steps))
    for mode in range(len(shape)):
        for _ in range(3):
            tb.unfolding_dot_khatri_rao(x, (weights, state.factors), mode)
        path_used = tb.last_kernel_path()
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
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if env.world == 1 and os.path.exists(tp):      # the ncu capture is of the 1-GPU launch of this workload
        try:
            traffic = json.load(open(tp)).get(key, {}).get(path_used)
        except Exception:
            traffic = None
    # context for the fraction: what a read-only LDG.128 pass over the same tensor (max |x|) streams on this box now
    read_only = None
    if dtype == torch.float32:
        try:
            for _ in range(2):
                tb.tensor_absmax(x)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                tb.tensor_absmax(x)
            e1.record()
            torch.cuda.synchronize()
            read_only = x.numel() * 4 / (e0.elapsed_time(e1) / 5 * 1e-3) / 1e9
        except Exception:
            read_only = None
    out["roofline"] = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
        "read_only_pass_gbs": read_only,
        "peak_source": peak_src, "kernel": f"MTTKRP ({path_used}), mean over the {len(shape)} modes, per call incl. "
                                           "its prep and split-K reduce launches",
        "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": avg_ms,
        "per_mode_gbs": [alg_bytes / (m * 1e-3) / 1e9 for m in mttkrp_ms], "frac_of_nominal_8TBs": achieved / 8000.0,
        "note": ("frac > 1: `peak` is the measured COPY bandwidth (read + write); a read-only stream like this one "
                 "can exceed it — see frac_of_nominal_8TBs") if achieved > peak else None}
    out["path"] = path_used

    # ---- launches per sweep (host-side count of one eager sweep) -------------------------
    c0 = tb.launch_count()
    state.sweep_eager(True)
    out["launches_per_sweep"] = int(tb.launch_count() - c0)

    # ---- timed region: K sweeps, inputs resident in HBM ----------------------------------
    mark0 = env.sampler.mark() if env.sampler else 0
    total_ms = timed_sweeps(env, state, args.steps, max(3, args.warmup))
    out["ms_per_step"] = total_ms / args.steps
    out["value"] = args.steps / (total_ms * 1e-3)
    out["rel_err"] = float(state.err[0].item())

    # ---- sustained: the same sweeps for >= ~2 s (clocks settle under the power cap) --------
    if not args.no_sustained:
        n_long = int(min(2000, max(args.steps, 2000.0 / max(out["ms_per_step"], 1e-3))))
        ms_long = timed_sweeps(env, state, n_long, 0)
        out["sustained"] = {"value": n_long / (ms_long * 1e-3), "unit": UNIT, "steps": n_long,
                            "ms_per_step": ms_long / n_long,
                            "what": "same sweeps timed over a >= 2 s window (SM clocks settled under the power cap)"}
    out["clock_marks"] = (mark0, env.sampler.mark() if env.sampler else 0)

    # ---- the same sweeps with one full MTTKRP per mode (N tensor passes, the reference's call pattern) ----
    if state.dimtree and headline:
        st3 = CPALS(x, weights, factors, comm=env.comm, shard_mode=0, dimtree=False)
        n3 = max(3, min(args.steps, 30))
        ms3 = timed_sweeps(env, st3, n3, 3)
        out["three_pass"] = {"value": n3 / (ms3 * 1e-3), "unit": UNIT, "ms_per_step": ms3 / n3, "steps": n3,
                             "final_rel_error": float(st3.err[0].item()),
                             "what": "same sweeps with a full MTTKRP per mode (no dimension-tree reuse): the call "
                                     "pattern SURVEY 8(d)'s bytes/sweep model describes"}
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
        out["ttm_pass"] = {"ms_per_launch": t_ms, "algorithmic_bytes_per_launch": t_bytes,
                           "achieved": t_bytes / (t_ms * 1e-3) / 1e9, "unit": "GB/s",
                           "kernel": f"mode_dot(X, F_last^T, last) ({tb.last_kernel_path()})"}
    return out


def e2e_leg(env, res, args):
    """The same sweep through the public API with HOST buffers: per step the pinned-host slab and factors go to the
    device, ||X||^2 + Grams + one full sweep run through tensorly_b200.CPALS, factors and error come back."""
    import torch
    from tensorly_b200.cp_als import CPALS
    x, state = res["x"], res["state"]
    dtype, esize = x.dtype, res["esize"]
    try:
        x_host = torch.empty(x.shape, dtype=dtype, pin_memory=True)
    except Exception as exc:
        return {"unavailable": f"cannot pin {x.numel() * esize / 2**30:.1f} GiB of host memory: {str(exc)[:120]}"}
    x_host.copy_(x)
    f_host = [torch.empty(f.shape, dtype=dtype, pin_memory=True).copy_(f) for f in state.factors]
    out_host = [torch.empty(f.shape, dtype=dtype, pin_memory=True) for f in state.factors]
    err_host = torch.empty(3, dtype=dtype, pin_memory=True)
    # the device copy of the step's input is the resident tensor's own buffer: every H2D copy overwrites it
    x_dev = x
    steps = max(2, min(args.steps, 6))

    def step():
        x_dev.copy_(x_host, non_blocking=True)                                 # H2D: the tensor slab
        fs = [h.to(env.device, non_blocking=True) for h in f_host]             # H2D: current factors
        st = CPALS(x_dev, res["weights"], fs, comm=env.comm, shard_mode=0)     # ||X||^2 + Grams
        st.sweep_eager(True)
        for o, f in zip(out_host, st.factors):
            o.copy_(f, non_blocking=True)                                      # D2H: updated factors
        err_host.copy_(st.err, non_blocking=True)                              # D2H: the error
        torch.cuda.current_stream().synchronize()

    for _ in range(2):
        step()
    env.barrier()
    t0 = time.perf_counter()
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    for _ in range(steps):
        step()
    a1.record()
    env.barrier()
    secs = env.max_over_ranks(max(a0.elapsed_time(a1) * 1e-3, time.perf_counter() - t0))
    R = state.rank
    fbytes = esize * R * sum(res["local_shape"])
    h2d = env.sum_over_ranks(esize * x.numel() + fbytes)
    d2h = env.sum_over_ranks(fbytes + 3 * esize)
    del x_host
    return {"value": steps / secs, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
            "steps": steps, "final_rel_error": float(err_host[0]),
            "note": "per step: pinned-host tensor slab + factors -> device, ||X||^2 + Grams + one full ALS sweep "
                    "through tensorly_b200.CPALS, factors + error -> host (PCIe-bound by construction)"}


def parity_leg(env):
    """Driver-visible correctness of the path that was just timed, on a shape small enough to run twice:
    the sharded trajectory (real kernels, real collectives) against the single-GPU one from the same tensor and
    the same initial factors, and the dimension-tree sweep against the N-pass sweep.  Gate: 1e-4 relative on the
    reconstruction errors (north star)."""
    import torch
    import tensorly_b200 as tb
    from tensorly_b200.cp_als import CPALS, _Comm
    shape, R, n = PARITY_SHAPE, PARITY_RANK, PARITY_SWEEPS
    x = device_slab(shape, 0, shape[0], torch.float32, env.device, seed=7, block=32)
    fs = initial_factors(shape, R, torch.float32, env.device, seed=8)
    w = torch.ones(R, device=env.device)
    solo = _Comm(None, sharded=False)

    def trajectory(st):
        errs = []
        for _ in range(n):
            st.sweep(True)
            errs.append(st.err[0].clone())
        return torch.stack(errs), st

    e_tree, st_tree = trajectory(CPALS(x, w, fs, comm=solo))
    e_npass, _ = trajectory(CPALS(x, w, fs, comm=solo, dimtree=False))
    dev_tree = float(torch.max(torch.abs(e_tree - e_npass) / e_npass))
    out = {"shape": list(shape), "rank": R, "sweeps": n, "gate": 1e-4,
           "dimension_tree_vs_n_pass_max_rel_err_dev": dev_tree}
    ok = dev_tree <= 1e-4
    if env.world > 1:
        lo, hi = tb.shard_bounds(shape[0], env.world, env.rank)
        fl = [f.clone() for f in fs]
        fl[0] = fl[0][lo:hi].contiguous()
        e_sh, st_sh = trajectory(CPALS(x[lo:hi].contiguous(), w, fl, comm=env.comm, shard_mode=0))
        dev_sh = float(torch.max(torch.abs(e_sh - e_tree) / e_tree))
        fdev = 0.0
        for k, (a, b) in enumerate(zip(st_sh.factors, st_tree.factors)):
            b = b[lo:hi] if k == 0 else b
            fdev = max(fdev, float(torch.linalg.norm(a - b) / torch.linalg.norm(b)))
        dev_sh = env.max_over_ranks(dev_sh)
        fdev = env.max_over_ranks(fdev)
        out["sharded_vs_single_gpu_max_rel_err_dev"] = dev_sh
        out["sharded_vs_single_gpu_max_factor_dev"] = fdev
        out["collective"] = env.comm_kind
        ok = ok and dev_sh <= 1e-4 and fdev <= 1e-2
        st_sh._graph = None
    out["ok"] = bool(ok)
    out["final_rel_error"] = float(e_tree[-1])
    return out


def fp64_leg(env):
    """fp64 MTTKRP / TTM on the SIMT DFMA path (B200's FP64 tensor and vector peaks are the same ~37 TFLOP/s, so
    DFMA is the fp64 roofline): 512^3, rank 32; intensity 2R flop / 8 B = 8 flop/B => compute-bound."""
    import torch
    import tensorly_b200 as tb
    shape, R = (512, 512, 512), 32
    g = torch.Generator(device=env.device).manual_seed(5)
    x = torch.rand(shape, generator=g, dtype=torch.float64, device=env.device)
    fs = [torch.rand((s, R), generator=g, dtype=torch.float64, device=env.device) for s in shape]
    res = {"shape": list(shape), "rank": R}
    flops = 2.0 * R * x.numel()
    nbytes = 8.0 * (x.numel() + R * sum(shape))

    def timeit(fn, reps=5):
        for _ in range(2):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps
    ms = [timeit(lambda m=m: tb.unfolding_dot_khatri_rao(x, (None, fs), m)) for m in range(3)]
    # DFMA peak of this GPU: 148 SMs x 64 DFMA/clk x 2 x SM clock
    peak_tf = 148 * 64 * 2 * 1.965e9 / 1e12
    res["mttkrp_ms_per_mode"] = ms
    res["mttkrp_tflops"] = [flops / (m * 1e-3) / 1e12 for m in ms]
    res["mttkrp_gbs"] = [nbytes / (m * 1e-3) / 1e9 for m in ms]
    res["fp64_peak_tflops"] = peak_tf
    res["mttkrp_frac_of_fp64_peak"] = statistics.mean(res["mttkrp_tflops"]) / peak_tf
    m_t = timeit(lambda: tb.mode_dot(x, fs[2], 2, transpose=True))
    res["ttm_ms"] = m_t
    res["ttm_tflops"] = flops / (m_t * 1e-3) / 1e12
    res["ttm_frac_of_fp64_peak"] = res["ttm_tflops"] / peak_tf
    res["kernel_path"] = tb.last_kernel_path()
    res["peak_source"] = "148 SMs x 64 DFMA/clk/SM x 2 flop x 1.965 GHz (nominal; FP64 tensor peak is the same on B200)"
    return res


def n4_leg(env):
    """SURVEY 8(f) n4 kernels on a C2-like problem (512 x 1024 x 1024 fp32, rank 32): reconstruction (writes the tensor
    once), the fused masked-ALS imputation pass (reads tensor + mask, writes tensor: 12 B per element), one masked
    sweep of the own driver, and one HALS mode update.  Both passes run on the tcgen05 kernel of recon_tc.cu."""
    import torch
    import tensorly_b200 as tb
    shape, R = (512, 1024, 1024), 32
    g = torch.Generator(device=env.device).manual_seed(9)
    x = torch.rand(shape, generator=g, device=env.device)
    fs = [torch.rand((s, R), generator=g, device=env.device) for s in shape]
    w = torch.ones(R, device=env.device)
    res = {"shape": list(shape), "rank": R}

    def timeit(fn, reps=5):
        for _ in range(2):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps
    n = x.numel()
    peak, _ = measured_peak_hbm()
    out = torch.empty_like(x)
    ms = timeit(lambda: tb.cp_to_tensor((w, fs), out=out))
    res["cp_to_tensor"] = {"ms": ms, "gbs_written": 4.0 * n / (ms * 1e-3) / 1e9, "tflops": 2.0 * R * n / (ms * 1e-3) / 1e12,
                           "frac_of_hbm_peak": 4.0 * n / (ms * 1e-3) / 1e9 / peak,
                           "kernel_path": tb.last_kernel_path(),
                           "bound": "HBM writes (tcgen05 kernel: 3xTF32 product, TMA-store epilogue)"}
    mask = (torch.rand(shape, generator=g, device=env.device) > 0.1).to(torch.float32)
    xi = x.clone()
    ms = timeit(lambda: tb.cp_impute(xi, mask, (w, fs), out=xi))
    res["cp_impute"] = {"ms": ms, "gbs": 12.0 * n / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": 12.0 * n / (ms * 1e-3) / 1e9 / peak,
                        "what": "x*mask + rec*(1-mask) in place + both norms: 12 B per element, rec never materialised"}
    # the same two passes at rank 64 (two 32-wide contraction chunks, one Khatri-Rao stage)
    fs64 = [torch.rand((s, 64), generator=g, device=env.device) for s in shape]
    w64 = torch.ones(64, device=env.device)
    ms = timeit(lambda: tb.cp_to_tensor((w64, fs64), out=out))
    res["cp_to_tensor_rank64"] = {"ms": ms, "gbs_written": 4.0 * n / (ms * 1e-3) / 1e9}
    ms = timeit(lambda: tb.cp_impute(xi, mask, (w64, fs64), out=xi))
    res["cp_impute_rank64"] = {"ms": ms, "gbs": 12.0 * n / (ms * 1e-3) / 1e9}
    del fs64
    st = tb.CPALS(x, w, fs, mask=mask)
    for _ in range(2):
        st.sweep(True)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(5):
        st.sweep(True)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    res["masked_sweep"] = {"ms": ms, "sweeps_per_s": 1e3 / ms, "what": "own driver with mask=: dimension-tree sweep + one imputation pass"}
    del st
    grams = [tb.gram(f) for f in fs]
    m = tb.unfolding_dot_khatri_rao(x, (None, fs), 0)
    f0 = fs[0].clone()
    ms = timeit(lambda: tb.hals_update(grams, 0, w, m, f0, n_iter_max=100, tol=1e-8), reps=3)
    res["hals_update"] = {"ms": ms, "rows": shape[0], "what": "one kernel = the whole hals_nnls call of a mode update (<= 100 passes over the rank)"}
    return res


def c4_leg(env, steps):
    """C4 (BASELINE configs[3]): non_negative_parafac (multiplicative updates) rank 64 on a random 256^4 fp32 tensor
    (17.2 GB): every mode's MTTKRP streams the tensor in place (mode 0 / middle / last views of a 4-way array; the
    reference would materialise a 4.3 GB Khatri-Rao matrix and a 17 GB permuted copy per mode)."""
    import torch
    import tensorly_b200 as tb
    shape, R = (256, 256, 256, 256), 64
    x = device_slab(shape, 0, shape[0], torch.float32, env.device, seed=4)
    g = torch.Generator(device=env.device).manual_seed(6)
    fs = [torch.rand((s, R), generator=g, device=env.device) for s in shape]
    w = torch.ones(R, device=env.device)
    out = {"workload": "C4: non_negative_parafac rank 64 on random 256x256x256x256 fp32", "unit": UNIT}
    st = tb.CPALS(x, w, fs, update="mu")
    for _ in range(3):
        st.sweep(True)
    n = max(5, min(steps, 20))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(n):
        st.sweep(True)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / n
    out.update({"value": 1e3 / ms, "ms_per_step": ms, "steps": n, "final_rel_error": float(st.err[0])})
    alg = 4.0 * (x.numel() + R * sum(shape))
    per_mode = []
    for mode in range(4):
        for _ in range(2):
            tb.unfolding_dot_khatri_rao(x, (None, st.factors), mode)
        torch.cuda.synchronize()
        a.record()
        for _ in range(3):
            tb.unfolding_dot_khatri_rao(x, (None, st.factors), mode)
        b.record()
        torch.cuda.synchronize()
        per_mode.append(alg / (a.elapsed_time(b) / 3 * 1e-3) / 1e9)
    peak, src = measured_peak_hbm()
    out["roofline"] = {"bound": "hbm", "achieved": statistics.mean(per_mode), "peak": peak, "unit": "GB/s",
                       "frac": statistics.mean(per_mode) / peak, "per_mode_gbs": per_mode, "kernel": f"MTTKRP ({tb.last_kernel_path()})",
                       "algorithmic_bytes_per_launch": alg, "peak_source": src}
    return out


def tucker_leg(env, steps):
    """C3 (BASELINE configs[2]): tucker HOOI rank [64,64,64] on random 512^3 fp32 — the TTM tensor-core path.
    Own driver (TTM chains + Gram + subspace iteration on the hand-written kernels) and, beside it, the unmodified
    reference driver on the b200 tenalg backend with torch's SVD and with the Gram+eigh plug-in.  Roofline of the
    dominant kernel = the first TTM of a chain (the only pass over the 537 MB tensor): HBM-bound at this intensity."""
    import torch
    import tensorly_b200 as tb
    shape, ranks = (512, 512, 512), [64, 64, 64]
    x = device_slab(shape, 0, shape[0], torch.float32, env.device, seed=3)
    out = {"workload": "C3: tucker HOOI rank [64,64,64] on random 512x512x512 fp32", "unit": UNIT}
    # own driver: sweeps timed with CUDA events after init + warm-up sweeps
    factors = tb.tucker_hooi._svd_init(tb.tucker_hooi.CudaOps, x, ranks, [0, 1, 2])
    st = tb.HOOI(x, ranks, [0, 1, 2], factors)
    c0 = tb.launch_count()
    st.sweep_eager()
    launches = int(tb.launch_count() - c0)
    for _ in range(3):
        st.sweep()
    n = max(5, min(steps, 20))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(n):
        st.sweep()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / n
    out.update({"value": 1e3 / ms, "ms_per_step": ms, "steps": n, "final_rel_error": float(st.err[0]),
                "launches_per_sweep": launches, "svd_iters": st.svd_iters,
                "what": "tensorly_b200.tucker sweep: 3 x (TTM chain skip=k, Gram of the unfolding, warm-started subspace "
                        "iteration) + core + error; no SVD/eigh, no host sync in the loop"})
    # the first TTM of a chain: X x_1 U1^T (the tensor pass), tcgen05 engine
    u = st.factors
    def ttm():
        return tb.mode_dot(x, u[1], 1, transpose=True)
    for _ in range(3):
        ttm()
    a.record()
    for _ in range(10):
        ttm()
    b.record()
    torch.cuda.synchronize()
    t_ms = a.elapsed_time(b) / 10
    nbytes = 4.0 * (x.numel() + x.numel() // 512 * 64 + 64 * 512)
    peak, peak_src = measured_peak_hbm()
    out["roofline"] = {"bound": "hbm", "kernel": f"mode_dot(X, U^T, mode 1) ({tb.last_kernel_path()}), first TTM of a HOOI chain",
                       "ms_per_launch": t_ms, "algorithmic_bytes_per_launch": nbytes, "achieved": nbytes / (t_ms * 1e-3) / 1e9,
                       "peak": peak, "unit": "GB/s", "frac": nbytes / (t_ms * 1e-3) / 1e9 / peak, "peak_source": peak_src,
                       "tflops_useful": 2.0 * 64 * x.numel() / (t_ms * 1e-3) / 1e12,
                       "note": "28 flop/B: 3xTF32 on tcgen05 keeps the pass HBM-bound; tensor-pipe utilisation in profiles/"}
    # unmodified reference driver on the backend
    try:
        tl = tb.import_tensorly()
        tl.set_backend("pytorch")
        tb.use()
        from tensorly.decomposition import tucker as ref_tucker

        def timed(k):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            _, errs = ref_tucker(x, ranks, n_iter_max=k, init="random", random_state=1, tol=0, return_errors=True)
            torch.cuda.synchronize()
            return time.perf_counter() - t0, float(errs[-1])
        timed(1)
        ta, _ = timed(1)
        tb_, e_ref = timed(6)
        out["reference_driver_on_b200"] = {"value": 5.0 / max(tb_ - ta, 1e-9), "unit": UNIT, "final_rel_error_6_sweeps": e_ref,
                                           "what": "tensorly.decomposition.tucker (unmodified, torch.linalg.svd) on tenalg 'b200'"}
        tb.use_gram_svd()
        try:
            timed(1)
            ta, _ = timed(1)
            tb_, _ = timed(6)
            out["reference_driver_on_b200"]["value_gram_svd"] = 5.0 / max(tb_ - ta, 1e-9)
        finally:
            tb.use_default_svd()
        # same init, same number of sweeps through the own driver.  The yardstick is exact HOOI with an fp64 SVD of the
        # projected unfolding (library SVD, checker only): the unmodified driver's fp32 cuSOLVER SVD returns factors
        # that are orthonormal only to ~1e-4, which inflates ||core|| and so UNDER-reports its error by ~7e-4.
        import numpy as np
        _, errs = tb.tucker(x, ranks, n_iter_max=6, init="random", random_state=1, tol=0, return_errors=True)
        rs = np.random.RandomState(1)
        rs.random_sample(ranks)
        fs = [torch.as_tensor(rs.random_sample((s_, r_))).to(env.device).float() for s_, r_ in zip(shape, ranks)]
        nx2 = float(tb.sumsq(x))
        exact = []
        for _ in range(6):
            for k in range(3):
                y = tb.multi_mode_dot(x, fs, skip=k, transpose=True)
                u, _, _ = torch.linalg.svd(tb.unfold(y, k, contiguous=True).double(), full_matrices=False)
                fs[k] = u[:, :ranks[k]].float().contiguous()
            core = tb.multi_mode_dot(x, fs, transpose=True)
            exact.append((abs(nx2 - float(tb.sumsq(core))) / nx2) ** 0.5)
        dev_exact = max(abs(a - b) / b for a, b in zip(errs, exact))
        out["parity_vs_exact_hooi_fp64_svd"] = {"own": errs, "exact": exact, "max_rel_dev": dev_exact, "gate": 1e-4,
                                                "ok": dev_exact <= 1e-4, "sweeps": 6, "init": "random, random_state=1"}
        out["reference_driver_on_b200"]["rel_dev_vs_exact_hooi_fp64_svd"] = abs(e_ref - exact[-1]) / exact[-1]
        out["parity_vs_reference_driver"] = {"own": errs[-1], "reference": e_ref, "rel_dev": abs(errs[-1] - e_ref) / e_ref,
                                             "note": "the reference figure comes from cuSOLVER's fp32 SVD; see parity_vs_exact_hooi_fp64_svd"}
    except Exception as exc:
        out["reference_driver_on_b200"] = {"unavailable": str(exc)[:200]}
    return out


def reference_driver_leg(env, key, res):
    """The UNMODIFIED tensorly.decomposition.parafac on the b200 tenalg backend (delta-iterations timing)."""
    import torch
    import tensorly_b200 as tb
    x, weights, factors = res["x"], res["weights"], res["factors"]
    R = WORKLOADS[key]["rank"]
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
        out = {"value": 10.0 / max(tb_ - ta, 1e-9), "unit": UNIT,
               "what": "tensorly.decomposition.parafac (unmodified) on tl.tenalg backend 'b200'"}
        tb.set_dimension_tree(True)
        try:
            timed(2)
            ta, tb_ = timed(2), timed(12)
            out["value_dimension_tree"] = 10.0 / max(tb_ - ta, 1e-9)
            out["what_dimension_tree"] = "same, with tensorly_b200.use(dimension_tree=True)"
        finally:
            tb.set_dimension_tree(False)
        if hasattr(tb, "use_fast_solve"):
            tb.use_fast_solve()
            try:
                tb.set_dimension_tree(True)
                timed(2)
                ta, tb_ = timed(2), timed(12)
                out["value_fast_solve"] = 10.0 / max(tb_ - ta, 1e-9)
                out["what_fast_solve"] = "same, with the dimension-tree cache and the sync-free small-system `solve` hook"
            finally:
                tb.set_dimension_tree(False)
                tb.use_default_solve()
        return out
    except Exception as exc:  # tensorly not importable on this box
        return {"unavailable": str(exc)[:200]}


def cpu_baseline_leg(key, budget_s, steps=2, warmup=1):
    """cpu_baseline of the `ours` line: the unmodified reference on the host cores, real shape when memory and
    the time budget allow (C2: yes), else a mode-0 slab, labelled."""
    wl = WORKLOADS[key]
    sample, frac, avail = reference_plan(wl["shape"], wl["rank"], wl["dtype"])
    if key == "c5" and sample[0] > 256:
        # ~30 s per real C5 sweep on 16 threads: keep the default bench run short, the full-shape timing is the
        # job of `--impl reference`.  One GPU's share at N = 8 (256 rows) is the bounded sample.
        sample, frac = (256,) + tuple(wl["shape"][1:]), 256 / wl["shape"][0]
    per, n, warm, kind, threads = cpu_reference_sweeps(sample, wl["rank"], wl["dtype"], steps, warmup, budget_s)
    what = (f"{n} timed sweeps (after {warm} warm-up) of unmodified tensorly parafac, numpy backend + core tenalg, "
            f"shape {tuple(sample)} rank {wl['rank']} {wl['dtype']}, {threads} threads")
    if frac != 1.0:
        what += f"; EXTRAPOLATED to {tuple(wl['shape'])} by the mode-0 row ratio {frac:.4g}"
    return {"value": frac / per, "unit": UNIT, "cores": threads, "kind": kind, "sample": what,
            "extrapolated": frac != 1.0, "measured_ms_per_sweep_at_sample": per * 1e3, "sample_shape": list(sample)}


def run_ours(args):
    import torch
    import torch.distributed as dist
    import tensorly_b200 as tb  # noqa: F401
    from tensorly_b200.cp_als import _Comm

    env = Env()
    env.world = int(os.environ.get("WORLD_SIZE", "1"))
    env.rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    env.device = torch.device("cuda", local)
    if env.world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=env.device)
    env.comm = _Comm(None, sharded=env.world > 1)
    env.comm_kind = None
    if env.world > 1:
        env.comm.all_reduce(torch.zeros(4, device=env.device))        # sets up the peer-memory path (or falls back)
        env.comm_kind = env.comm.kind

    def barrier():
        if env.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        t = torch.tensor([v], dtype=torch.float64, device=env.device)
        if env.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(v):
        t = torch.tensor([v], dtype=torch.float64, device=env.device)
        if env.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())
    env.barrier, env.max_over_ranks, env.sum_over_ranks = barrier, max_over_ranks, sum_over_ranks
    env.sampler = ClockSampler(local) if env.rank == 0 else None
    if env.sampler:
        env.sampler.start()

    def guarded(fn, *a):
        """Rank-0-only secondary blocks must never cost the headline line."""
        try:
            return fn(*a)
        except Exception as exc:
            return {"unavailable": f"{type(exc).__name__}: {str(exc)[:200]}"}

    key = args.workload
    res = bench_workload(env, key, args, headline=True)
    state = res["state"]
    clocks_timed = env.sampler.summary(*res["clock_marks"]) if env.sampler else None
    parity = parity_leg(env)
    ref_driver = None
    if env.rank == 0 and env.world == 1 and not args.no_refdriver:
        ref_driver = guarded(reference_driver_leg, env, key, res)
    e2e = None if args.no_e2e else e2e_leg(env, res, args)
    shape, R = WORKLOADS[key]["shape"], WORKLOADS[key]["rank"]
    dimtree = state.dimtree
    # free the headline tensor before the secondary block
    state._graph = None
    res.pop("state"), res.pop("x"), res.pop("factors")
    del state
    torch.cuda.empty_cache()

    c2 = None
    if key != "c2" and not args.no_c2:
        r2 = bench_workload(env, "c2", args, headline=False)
        c2 = {"workload": WORKLOADS["c2"]["name"], "value": r2["value"], "unit": UNIT, "ms_per_step": r2["ms_per_step"],
              "steps": args.steps, "sustained": r2.get("sustained"), "roofline": r2["roofline"],
              "final_rel_error": r2["rel_err"], "launches_per_sweep": r2["launches_per_sweep"],
              "clocks": env.sampler.summary(*r2["clock_marks"]) if env.sampler else None}
        if env.rank == 0 and env.world == 1 and not args.no_refdriver:
            c2["reference_driver_on_b200"] = guarded(reference_driver_leg, env, "c2", r2)
        r2["state"]._graph = None
        del r2
        torch.cuda.empty_cache()
    solo = env.rank == 0 and env.world == 1
    fp64 = guarded(fp64_leg, env) if (solo and not args.no_fp64) else None
    n4 = guarded(n4_leg, env) if (solo and not args.no_n4) else None
    c3 = guarded(tucker_leg, env, args.steps) if (solo and not args.no_c3) else None
    c4 = guarded(c4_leg, env, args.steps) if (solo and not args.no_c4) else None
    clocks = env.sampler.stop() if env.sampler else None

    # ---- CPU baselines on the host cores (rank 0, N=1 only), bounded samples ---------------
    cpu = None
    if env.rank == 0 and env.world == 1 and not args.no_cpu:
        cpu = guarded(cpu_baseline_leg, key, args.cpu_budget_s)
        if c2 is not None:
            c2["cpu_baseline"] = guarded(cpu_baseline_leg, "c2", args.cpu_budget_s)

    if env.rank == 0:
        wl = WORKLOADS[key]
        if clocks_timed is not None and clocks is not None:
            clocks_timed["whole_run"] = clocks
        line = {
            "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": env.world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": res["ms_per_step"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32" if wl["dtype"] == "float32" else "f64",
            "data": "synthetic",
            "config": {"workload": wl["name"], "shape": list(shape), "rank": R,
                       "sharding": f"mode-0 slabs over {env.world} GPU(s), the same seeded tensor at every N",
                       "l2": "inputs larger than L2 (tensor slab %.2f GB per GPU >> 126 MB)"
                             % (res["esize"] * (shape[0] // env.world) * shape[1] * shape[2] / 1e9),
                       "kernel_path": res["path"],
                       "arithmetic": ("fp32 in / fp32 out; products on tcgen05 as an error-compensated split (tcgen05-f16: fp16 "
                                      "hi/lo x 3 products on x * 2^k with max |x| from one pass at set-up, tcgen05: 3xTF32), fp32 "
                                      "accumulation drained to registers every <= 16 units; gate 1e-5 vs fp64 in the tests"),
                       "collective": env.comm_kind,
                       "sweep": ("dimension-tree ALS sweep: T = X x_last F_last^T (one tensor pass) -> MTTKRP of every "
                                 "earlier mode from T, full MTTKRP for the last mode (second tensor pass); identical "
                                 "factor updates, parity-tested against the N-pass sweep" if dimtree else
                                 "one full MTTKRP per mode") + " + Gram-Hadamard LU solve + Gram per mode + error"},
            "roofline": res["roofline"],
            "sustained": res.get("sustained"),
            "three_pass": res.get("three_pass"),
            "ttm_pass": res.get("ttm_pass"),
            "cpu_baseline": cpu,
            "e2e": e2e,
            "parity": parity,
            "gpu_launches": int(res["launches_per_sweep"] * args.steps),
            "launches_per_sweep": res["launches_per_sweep"],
            "clocks": clocks_timed,
            "final_rel_error": res["rel_err"],
            "reference_driver_on_b200": ref_driver,
            "c2": c2,
            "c3": c3,
            "c4": c4,
            "fp64": fp64,
            "n4": n4,
        }
        print(json.dumps(line), flush=True)
    if env.world > 1:
        # teardown must never turn a finished measurement into a hang: drop the CUDA graphs before the
        # communicator, and bail out hard if the collective teardown stalls anyway
        import gc

        def _bail():
            time.sleep(30)
            os._exit(0)
        threading.Thread(target=_bail, daemon=True).start()
        gc.collect()
        torch.cuda.synchronize()
        dist.barrier()
        env.comm.close()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c5", choices=sorted(WORKLOADS))
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-c2", action="store_true")
    ap.add_argument("--no-c4", action="store_true", help="skip the 256^4 non-negative CP block")
    ap.add_argument("--no-fp64", action="store_true")
    ap.add_argument("--no-n4", action="store_true", help="skip the reconstruction / imputation / HALS block")
    ap.add_argument("--no-c3", action="store_true")
    ap.add_argument("--no-sustained", action="store_true")
    ap.add_argument("--no-refdriver", action="store_true")
    ap.add_argument("--ref-budget-s", type=float, default=180.0)
    ap.add_argument("--cpu-budget-s", type=float, default=30.0)
    args = ap.parse_args()
    if args.impl == "reference":
        # torchrun exports OMP_NUM_THREADS=1: undo it before numpy/OpenBLAS load, the CPU arm gets every host thread
        n = str(host_threads())
        for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
            os.environ[var] = n
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
