#!/usr/bin/env python
"""Benchmark of the hot path named by BASELINE.json: SSFM throughput in Msamples/s (2-pol, per
span-step) on the cfg2 workload (11-ch WDM-like 2^20-sample dual-pol waveform, 10 x 80 km spans,
hz = 0.08 km -> 1001 executed loop steps per span, fixed step, EDFA per span).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One bench "step" = one full propagation of one waveform (10 spans) per GPU.
  value : N_samples x executed SSFM loop steps x n_gpus / time, input planar rows resident in HBM
  e2e   : the same metric through the public drop-in call manakovSSF(numpy, param) -> numpy
          (pinned host input, H2D + pack + run + unpack + D2H inside the timed region)
  roofline     : the fused nonlinear iteration kernel, timed in situ with CUDA events
  cpu_baseline : the CPU oracle port (numpy restatement of the reference) on a bounded sample
Under torchrun every rank propagates its own independent realisation (weak scaling) and the final
fields are gathered with one NCCL all_gather; time = max over ranks.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_SAMPLES = 1 << 20
FS = 512e9            # 32 GBd x 16 SpS
N_CH, CH_SPACING = 11, 37.5e9
P_CH_W = 10 ** (-2 / 10) * 1e-3  # -2 dBm per channel
WORKLOAD = "cfg2: 11-ch WDM-like DP waveform, 2^20 samples, 10x80 km, hz=0.08 km (1001 steps/span), fixed step, amp=edfa"


class Bag:
    def __init__(self, **kw):
        self.__dict__.update(kw)


def channel_param(n_spans=10, **over):
    p = Bag(Fs=FS, Ltotal=80 * n_spans, Lspan=80, hz=0.08, alpha=0.2, D=16, gamma=1.3, Fc=193.1e12,
            amp="edfa", NF=4.5, maxIter=10, tol=1e-5, nlprMethod=False, maxNlinPhaseRot=2e-2, seed=456,
            prgsBar=False, saveSpanN=[], returnParameters=False, prec=np.complex128)
    p.__dict__.update(over)
    return p


def synth_waveform(seed, n=N_SAMPLES, dtype=np.complex128):
    """Band-limited complex Gaussian (N, 2) occupying the 11 x 37.5 GHz WDM band, 11 x -2 dBm."""
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(n, 2)) + 1j * rng.normal(size=(n, 2))
    f = np.fft.fftfreq(n) * FS
    X[np.abs(f) > N_CH * CH_SPACING / 2] = 0
    x = np.fft.ifft(X, axis=0)
    x *= np.sqrt(N_CH * P_CH_W / np.mean(np.sum(np.abs(x) ** 2, axis=1)))
    return np.ascontiguousarray(x.astype(dtype))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(engine):
    """DRAM bytes per launch of the roofline kernel from the committed ncu --set full capture, with its provenance (the
    figure is a profile constant, not measured in this run)."""
    for name in ("r2_traffic.json", "r1_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                t = json.load(f)
            v = t["k_time_iter_bytes"] if engine == "fused" else t["k_manakov_nl_iter_bytes"]
            return v, {"file": f"profiles/{name}", "source": t.get("source"), "captured": t.get("captured")}
        except Exception:
            continue
    return None, None


def cpu_oracle_rate(n_steps, n=N_SAMPLES, seed=0):
    """Time the CPU oracle port on a bounded sample: n_steps fixed steps of the cfg2 fiber at full N."""
    from oracle import fiber_oracle as fo
    x = synth_waveform(seed, n)
    cfg = fo.FiberConfig(Fs=FS, Ltotal=0.08 * n_steps, Lspan=0.08 * n_steps, hz=0.08, amp=None, nlprMethod=False)
    st = {}
    fo.manakov(x[:4096], fo.FiberConfig(Fs=FS, Ltotal=0.08, Lspan=0.08, hz=0.08, amp=None, nlprMethod=False))  # warm-up
    t0 = time.perf_counter()
    fo.manakov(x, cfg, stats=st)
    dt = time.perf_counter() - t0
    return n * st["steps"] / dt / 1e6, st["steps"], st["iterations"], dt


def nl_pass_microbench(torch, lib, _cabi, n=N_SAMPLES, reps=60):
    """The standalone fused nonlinear-step kernel (k_manakov_nl<false,2>: convergence sums + Kerr phase +
    rotation, 68 algorithmic bytes per 2-pol sample) timed alone with CUDA events.  Three buffer sets
    (3 x 68 MB > L2) are cycled so that every launch streams from HBM."""
    sets = []
    for i in range(3):
        g = torch.Generator(device="cuda").manual_seed(i)
        mk = lambda: (torch.randn((2, n, 2), device="cuda", generator=g) * 0.03).contiguous()
        sets.append(dict(ehd=mk(), efd=mk(), ec=mk(), pch=torch.rand((1, n), device="cuda", generator=g) * 1e-3,
                         out=torch.empty((2, n, 2), device="cuda"), sums=torch.zeros(3, dtype=torch.float64, device="cuda")))
    vp = C.c_void_p
    st = vp(_cabi.stream_ptr(torch))

    def launch(b):
        _cabi.check(lib.ocb_manakov_nl_pass(vp(b["ehd"].data_ptr()), vp(b["efd"].data_ptr()), vp(b["ec"].data_ptr()),
                                            vp(b["pch"].data_ptr()), vp(b["out"].data_ptr()), vp(b["sums"].data_ptr()),
                                            n, 1, 1.3, 0.08, 1, st), "nl_pass")
    for i in range(6):
        launch(sets[i % 3])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        launch(sets[i % 3])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3  # us per launch


def cupy_standin_rate(torch, rows0, prm, n_steps=40):
    """Stand-in for the reference's CuPy path (optic/models/modelsGPU.py:428-482), which cannot run here
    (no cupy): the same loop op for op with torch.fft and separate elementwise torch ops — operator
    rebuilt every step, four copies per iteration, host sync on `lim < tol` — on the same B200, complex64.
    Reported for context only; it is not the product path."""
    import math
    Ex, Ey = rows0[0:1].clone(), rows0[1:2].clone()
    n = Ex.shape[1]
    alpha = prm.alpha / (10 * math.log10(math.e))
    lam = 299792.458 / prm.Fc
    beta2 = -(prm.D * lam**2) / (2 * math.pi * 299792.458)
    w = 2 * math.pi * prm.Fs * torch.fft.fftfreq(n, device="cuda", dtype=torch.float64)
    arg = (-(alpha / 2) + 1j * (beta2 / 2) * w**2).to(torch.complex64).reshape(1, -1)
    g = prm.gamma
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    Exc, Eyc = Ex.clone(), Ey.clone()
    iters = 0
    for _ in range(n_steps):
        Pch = Ex * torch.conj(Ex) + Ey * torch.conj(Ey)
        phi = ((8 / 9) * g * (Pch + Exc * torch.conj(Exc) + Eyc * torch.conj(Eyc)) / 2).real
        hz = prm.hz
        lin = torch.exp(arg * (hz / 2))
        Exh = torch.fft.ifft(torch.fft.fft(Ex) * lin)
        Eyh = torch.fft.ifft(torch.fft.fft(Ey) * lin)
        for it in range(prm.maxIter):
            rot = torch.exp(1j * phi * hz)
            Exf = torch.fft.ifft(torch.fft.fft(Exh * rot) * lin)
            Eyf = torch.fft.ifft(torch.fft.fft(Eyh * rot) * lin)
            lim = torch.sqrt(torch.linalg.norm(Exf - Exc) ** 2 + torch.linalg.norm(Eyf - Eyc) ** 2) / \
                torch.sqrt(torch.linalg.norm(Exc) ** 2 + torch.linalg.norm(Eyc) ** 2)
            Exc, Eyc = Exf.clone(), Eyf.clone()
            iters += 1
            if float(lim) < prm.tol:  # device -> host sync, like the CuPy `if lim < tol`
                break
            phi = ((8 / 9) * g * (Pch + Exc * torch.conj(Exc) + Eyc * torch.conj(Eyc)) / 2).real
        Ex, Ey = Exf.clone(), Eyf.clone()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return n * n_steps / dt / 1e6, iters / n_steps


REF_STUBS = ["matplotlib", "matplotlib.pyplot", "matplotlib.mlab", "matplotlib.cm", "matplotlib.colors",
             "matplotlib.animation", "mpl_scatter_density", "simple_pid", "prettytable"]


def import_reference():
    """The UNMODIFIED reference's manakovSSF (OptiCommPy v0.11.0): baseline/_ref (pip --target install of
    /root/reference, travels to the GPU box), else /root/reference.  Plot-only dependencies are stubbed and numba's
    cache is pointed at a scratch directory.  Returns (manakovSSF, parameters, where) or None."""
    from unittest.mock import MagicMock
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/ocb_numba_cache")
    sys.dont_write_bytecode = True
    for cand in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if not os.path.isdir(os.path.join(cand, "optic")):
            continue
        for m in REF_STUBS:
            sys.modules.setdefault(m, MagicMock())
        sys.path.insert(0, cand)
        try:
            from optic.models.channels import manakovSSF
            from optic.utils import parameters
            return manakovSSF, parameters, cand
        except Exception:
            sys.path.remove(cand)
            for k in [k for k in sys.modules if k == "optic" or k.startswith("optic.")]:
                del sys.modules[k]
    return None


def reference_rate(n_steps, n=N_SAMPLES, seed=0):
    """Time the reference's own manakovSSF (CPU, complex128, single thread like the reference) on a bounded sample of
    the bench workload: n_steps fixed steps of the cfg2 fiber at full N.  Falls back to the oracle port if the
    reference cannot be imported.  Returns (Msamples/s, steps, seconds, kind)."""
    ref = import_reference()
    x = synth_waveform(seed, n)
    if ref is None:
        rate, st, _, dt = cpu_oracle_rate(n_steps, n, seed)
        return rate, st, dt, "port"
    manakovSSF, parameters, _ = ref

    def prm(steps):
        p = parameters()
        p.Fs, p.Ltotal, p.Lspan, p.hz = FS, 0.08 * steps, 0.08 * steps, 0.08
        p.alpha, p.D, p.gamma, p.Fc = 0.2, 16, 1.3, 193.1e12
        p.amp, p.nlprMethod, p.maxIter, p.tol = None, False, 10, 1e-5
        p.prgsBar, p.saveSpanN = False, []
        return p
    manakovSSF(x[:4096], prm(1))  # warm-up: numba JIT of nlinPhaseRot, FFT plan caches
    t0 = time.perf_counter()
    manakovSSF(x, prm(n_steps))
    dt = time.perf_counter() - t0
    # executed loop steps: the reference's `while z < Lspan` loop (channels.py:387-441), restated on the host
    z, steps = 0.0, 0
    while z < 0.08 * n_steps:
        z += 0.08 if not (0.08 * n_steps - z < 0.08) else (0.08 * n_steps - z)
        steps += 1
    return n * steps / dt / 1e6, steps, dt, "reference"


def _ref_worker(args):
    n_steps, seed = args
    os.environ["OMP_NUM_THREADS"] = "1"
    return reference_rate(n_steps, seed=seed)


def static_config(spans, world):
    """The workload description shared by both arms (nothing measured in here, so the two lines carry the same dict)."""
    return {"workload": WORKLOAD, "n_samples": N_SAMPLES, "spans": spans, "Lspan_km": 80, "hz_km": 0.08, "Fs_GSa": FS / 1e9,
            "steps_per_span_executed": 1001, "l2": "256 MiB flush write between timed steps",
            "parallelism": f"{world} independent realisation(s), one per GPU, NCCL all_gather of the final field"}


def run_reference(args):
    """--impl reference: the reference's own CPU manakovSSF (unmodified OptiCommPy from baseline/_ref) on all host
    cores, one independent realisation per process (the reference is single-threaded), each bench step a bounded
    sample of the workload.  The pool is created and warmed once, so every timed step sees JIT-compiled code."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = min(os.cpu_count() or 1, 32)
    steps_per_sample = 2
    times, kind = [], "reference"
    with mp.get_context("spawn").Pool(cores) as pool:
        for it in range(max(1, args.warmup) + args.steps):
            res = pool.map(_ref_worker, [(steps_per_sample, 100 * it + i) for i in range(cores)], chunksize=1)
            dt = max(r[2] for r in res)  # slowest worker's propagation time (input synthesis and JIT excluded)
            kind = res[0][3]
            if it >= max(1, args.warmup):
                times.append((dt, sum(r[1] for r in res)))
    tot_t = sum(t for t, _ in times)
    tot_steps = sum(s for _, s in times)
    value = N_SAMPLES * tot_steps / tot_t / 1e6
    sample = (f"{steps_per_sample} fixed SSFM steps (hz = 0.08 km) of the cfg2 fiber at N=2^20 per process x {cores} "
              f"processes per bench step, {'optic.models.channels.manakovSSF (unmodified reference)' if kind == 'reference' else 'oracle port'}")
    print(json.dumps({
        "impl": "reference", "metric": "SSFM Msamples/s (2-pol, per span-step)", "value": value, "unit": "Msamples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(1, len(times)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "c128 (f64 pairs)", "data": "synthetic",
        "config": static_config(args.spans, int(os.environ.get("WORLD_SIZE", "1"))),
        "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--spans", type=int, default=10, help="spans per propagation (10 = cfg2)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the extras: NL-kernel microbench, CuPy stand-in, cfg1, rx_chain (cfg3), cfg4_dbp, cfg5_mc")
    ap.add_argument("--nl-only", action="store_true", help="only time the standalone nonlinear-step kernel")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from opticommpy_b200 import _cabi, _engine
    from opticommpy_b200.channels import manakov_rows_device, manakovSSF

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # the version banner goes to stdout, where exactly one JSON line belongs
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _cabi.lib()
    st = C.c_void_p(_cabi.stream_ptr(torch))
    if args.nl_only:
        us = nl_pass_microbench(torch, lib, _cabi)
        print(json.dumps({"nl_pass_us": us, "GBps": 68.0 * N_SAMPLES / (us * 1e-6) / 1e9,
                          "frac_of_measured_peak": 68.0 * N_SAMPLES / (us * 1e-6) / 1e9 / measured_hbm_peak()[0]}))
        return

    prm = channel_param(args.spans)
    # ---- inputs: one independent realisation per rank, pinned on the host, pristine copy in HBM
    host = torch.empty((N_SAMPLES, 2), dtype=torch.complex128, pin_memory=True)
    host_np = host.numpy()
    host_np[...] = synth_waveform(1000 + rank)
    d_raw = host.to("cuda", non_blocking=False)
    rows0 = torch.empty((2, N_SAMPLES), dtype=torch.complex64, device="cuda")
    _cabi.check(lib.ocb_pack_fields(C.c_void_p(d_raw.data_ptr()), _cabi.OCB_C128, N_SAMPLES, 2, 1,
                                    C.c_void_p(rows0.data_ptr()), st), "pack")
    rows = torch.empty_like(rows0)
    gathered = [torch.empty_like(rows0) for _ in range(world)] if world > 1 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def one_step():
        rows.copy_(rows0)
        s = manakov_rows_device(rows, prm, +1)
        if world > 1:
            dist.all_gather(gathered, rows)  # the path's only collective: final gather over NVLink
        return s

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        stats = one_step()
    barrier()
    plan = _engine.get_plan(N_SAMPLES, 2)

    # ---- timed region: device-resident
    _cabi.launch_count_reset()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot_ms, tot_steps, tot_iters = 0.0, 0, 0
    with ClockSampler(local) as clk:
        for _ in range(args.steps):
            flush.zero_()  # L2 flush between timed iterations
            barrier()
            ev0.record()
            stats = one_step()
            ev1.record()
            barrier()
            tot_ms += ev0.elapsed_time(ev1)
            tot_steps += stats["steps"]
            tot_iters += stats["iterations"]
    launches = _cabi.launch_count()
    # ---- in-situ kernel timing: one more span of the same propagation with CUDA events recorded around
    # every launch of the step loop (2048 samples per kernel kind); kept out of `value` because the
    # event pairs serialise the otherwise speculative launch stream
    prof = (C.c_double * 6)()
    pass_us = {}
    if rank == 0:
        _cabi.check(lib.ocb_ssfm_plan_profile(plan.handle, 1), "profile on")
        rows.copy_(rows0)
        manakov_rows_device(rows, channel_param(1), +1)
        _cabi.check(lib.ocb_ssfm_plan_profile_read(plan.handle, prof), "profile read")
        _cabi.check(lib.ocb_ssfm_plan_profile(plan.handle, 0), "profile off")
        # Per-pass launch duration as the step loop sees it: 200 back-to-back launches of each pass kernel on the
        # plan's own buffers (L2-resident like in the loop, programmatic dependent launch on), one CUDA-event pair
        # around the batch on the launching stream.  The event pairs above bracket single launches instead, which
        # adds the event overhead and removes the launch overlap, so they are reported as a secondary figure.
        if plan.engine == "fused":
            for which, name in enumerate(["k_freq", "k_time_FIRST", "k_time_ITER", "k_time_ITERF", "k_time_FWD"]):
                us = C.c_double()
                _cabi.check(lib.ocb_ssfm_plan_pass_time(plan.handle, which, 200, C.byref(us), st), "pass_time")
                pass_us[name] = us.value
    t = torch.tensor([tot_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    tot_ms = float(t.item())
    value = N_SAMPLES * tot_steps * world / (tot_ms * 1e-3) / 1e6

    # ---- e2e: public drop-in API, numpy (pinned) in -> numpy out
    e2e_t, e2e_steps = 0.0, 0
    for i in range(1 + args.steps):
        p2 = channel_param(args.spans)
        barrier()
        t0 = time.perf_counter()
        out = manakovSSF(host_np, p2)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if i > 0:  # first call warms the staging allocations
            e2e_t += dt
            e2e_steps += p2._b200_stats["steps"]
    te = torch.tensor([e2e_t], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = N_SAMPLES * e2e_steps * world / float(te.item()) / 1e6
    noise_bytes = N_SAMPLES * 8  # injected ASE realisation (seeded reference mode)

    sharded = {}
    if not args.no_extras:
        import bench_extras as bx
        for name, fn in (("cfg4_dbp", bx.extra_cfg4_dbp), ("cfg5_mc", bx.extra_cfg5_mc)):
            try:
                sharded[name] = fn(torch, dist, world, rank)
            except Exception as e:
                sharded[name] = {"error": f"{type(e).__name__}: {e}"}
                if world > 1:
                    raise  # a rank that drops out of a collective would hang the others
    if rank == 0:
        peak, peak_src = measured_hbm_peak()
        mean_I = tot_iters / max(1, tot_steps)
        nl_ms, nl_n = prof[0], prof[1]
        # fused engine: W row 16 + Ec 16 + Ehd 16 + Pch 4 read, new iterate 16 + W row 16 written = 84 B
        # cuFFT engine NL pass: Efd 16 + Ec 16 + Ehd 16 + Pch 4 read, rotated field 16 written = 68 B
        bytes_nl = (84.0 if plan.engine == "fused" else 68.0) * N_SAMPLES
        ev_us = 1e3 * nl_ms / max(nl_n, 1) if nl_n else None
        kern_us = pass_us.get("k_time_ITER", ev_us)
        ach = bytes_nl / (kern_us * 1e-6) / 1e9 if kern_us else None
        step_bytes = (64.0 + 84.0 * mean_I) * N_SAMPLES  # SURVEY §8d model per SSFM step
        step_ach = step_bytes * tot_steps / (tot_ms * 1e-3) / 1e9
        line = {
            "metric": "SSFM Msamples/s (2-pol, per span-step)", "value": value, "unit": "Msamples/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "c64 (f32 pairs; f64 linear-operator phase)",
            "data": "synthetic",
            "config": static_config(args.spans, world),
            "measured": {"ssfm_steps_per_bench_step": tot_steps // args.steps, "mean_fixed_point_iterations": mean_I},
            "clocks": clk.summary(),
            "e2e": {"value": e2e_value, "unit": "Msamples/s", "h2d_bytes_per_step": int(host_np.nbytes + noise_bytes),
                    "d2h_bytes_per_step": int(out.nbytes)},
            "gpu_launches": int(launches),
            "engine": plan.engine,
            "roofline": {"bound": "hbm", "kernel": ("k_time<32,2,TM_ITER> (IFFT_N1 + convergence sums + Kerr phase/rotation + FFT_N1)"
                                                    if plan.engine == "fused" else
                                                    "k_manakov_nl<false,2> (convergence sums + Kerr phase/rotation)"),
                         "achieved": ach, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                         "frac": (ach / peak) if ach else None, "traffic": ncu_traffic(plan.engine)[0],
                         "traffic_source": ncu_traffic(plan.engine)[1],
                         "launches_timed": 200 if pass_us else int(nl_n), "avg_us": kern_us,
                         "how": ("200 back-to-back launches on the plan's L2-resident buffers, one CUDA-event pair on the launch stream"
                                 if pass_us else "CUDA-event pair around every launch of one extra span"),
                         "avg_us_single_launch_event_pairs": ev_us, "bytes_per_launch": bytes_nl},
            "roofline_step": {"model": "64 + 84*I bytes per 2-pol sample-step", "achieved": step_ach, "peak": peak,
                              "unit": "GB/s", "frac": step_ach / peak,
                              "linear_half_step_avg_us": 1e3 * prof[4] / max(prof[5], 1),
                              "nl_first_avg_us": 1e3 * prof[2] / max(prof[3], 1),
                              "pass_us": pass_us,
                              "pass_bytes_per_sample": {"k_freq": 40, "k_time_FIRST": 68, "k_time_ITER": 84,
                                                        "k_time_ITERF": 64, "k_time_FWD": 32}},
        }
        if not args.no_extras:
            us = nl_pass_microbench(torch, lib, _cabi)
            a_nl = 68.0 * N_SAMPLES / (us * 1e-6) / 1e9
            line["roofline_nl_step"] = {"kernel": "k_manakov_nl<false,2> (standalone fused nonlinear step, ocb_manakov_nl_pass)",
                                        "bound": "hbm", "bytes_per_launch": 68.0 * N_SAMPLES, "avg_us": us, "achieved": a_nl,
                                        "peak": peak, "unit": "GB/s", "frac": a_nl / peak,
                                        "note": "3 rotating buffer sets (204 MB > L2), 60 launches, CUDA events"}
            sr, si = cupy_standin_rate(torch, rows0, channel_param(1))
            line["cupy_standin"] = {"value": sr, "unit": "Msamples/s", "mean_iterations": si,
                                    "what": "op-for-op torch.fft restatement of optic/models/modelsGPU.py:428-482 (unfused, host sync per iteration), 40 steps, complex64, same B200"}
        line.update(sharded)
        if not args.no_extras and world == 1:
            import bench_extras as bx
            for name, fn in (("cfg1", lambda: bx.extra_cfg1(torch)), ("rx_chain", lambda: bx.extra_rx_chain(torch, peak)),
                             ("cfg2_concurrent", lambda: bx.extra_cfg2_concurrent(torch, rows0)),
                             ("wdm_tx", lambda: bx.extra_wdm_tx(torch))):
                try:
                    line[name] = fn()
                except Exception as e:  # an extra must never take the headline down
                    line[name] = {"error": f"{type(e).__name__}: {e}"}
        if not args.no_cpu_baseline:
            rate, s_, dt, kind = reference_rate(4)
            line["cpu_baseline"] = {"value": rate, "unit": "Msamples/s", "cores": 1, "kind": kind,
                                    "sample": f"{s_} fixed SSFM steps of the same fiber at N=2^20 in {dt:.1f} s, "
                                              + ("optic.models.channels.manakovSSF of the unmodified reference (baseline/_ref), complex128"
                                                 if kind == "reference" else "numpy oracle port")}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
