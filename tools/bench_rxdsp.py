"""cfg3 (BASELINE configs[2]): full Rx DSP chain edc -> 2x2 mimoAdaptEqualizer (CMA->RDE, 31 taps) -> bps
(64 test phases, window 25) on 2^22 input samples (2^21 symbols x 2 pols at 2 SpS).  Prints one JSON
line with per-stage and chain throughput through the public drop-in calls (numpy in, numpy out), next
to the CPU oracle timed on a bounded sample.  Not the headline metric; kept under tools/."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


class Bag:
    def __init__(self, **kw):
        self.__dict__.update(kw)


def make_signal(nsym, seed=0):
    from opticommpy_b200.modulation import grayMapping
    from oracle import fiber_oracle as fo
    rng = np.random.default_rng(seed)
    c = grayMapping(16, "qam").astype(np.complex128)
    c /= np.sqrt(np.mean(np.abs(c) ** 2))
    sym = c[rng.integers(0, 16, size=(nsym, 2))]
    up = np.zeros((2 * nsym, 2), dtype=complex)
    up[0::2] = sym
    # root-raised-cosine-like low-pass (2 SpS), 2x2 rotation, 800 km of CD, phase noise, AWGN
    X = np.fft.fft(up, axis=0)
    f = np.fft.fftfreq(2 * nsym)
    X *= (np.abs(f) < 0.27)[:, None]
    x = np.fft.ifft(X, axis=0) * 2
    th = 0.5
    rot = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    x = x @ rot.T
    x = fo.linear_fiber(x, 800, 0.0, 16, 193.1e12, 64e9)
    pn = np.cumsum(rng.normal(scale=np.sqrt(2 * np.pi * 100e3 / 64e9), size=2 * nsym))
    x = x * np.exp(1j * pn)[:, None]
    x += 0.07 * (rng.normal(size=x.shape) + 1j * rng.normal(size=x.shape))
    return x / np.sqrt(np.mean(np.abs(x) ** 2)), sym


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nsym-log2", type=int, default=21)
    ap.add_argument("--cpu-nsym-log2", type=int, default=16)
    a = ap.parse_args()
    import torch
    from opticommpy_b200.carrierRecovery import cpr
    from opticommpy_b200.equalization import edc, mimoAdaptEqualizer
    from oracle import rxdsp_oracle as ro
    from opticommpy_b200.modulation import grayMapping

    nsym = 1 << a.nsym_log2
    x, sym = make_signal(nsym)
    pe = Bag(L=800, D=16, Fc=193.1e12, Fs=64e9, Rs=32e9)
    pq = Bag(nTaps=31, SpS=2, M=16, constType="qam", alg=["cma", "rde"], mu=[5e-3, 2e-4],
             L=[int(0.2 * nsym), int(0.8 * nsym)], prgsBar=False)
    pc = Bag(alg="bps", M=16, constType="qam", N=25, B=64, runFOE=False)

    def timed(fn, reps=3):
        fn()  # warm-up
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
        return float(np.median(ts)), out

    t_edc, y1 = timed(lambda: edc(x, pe))
    t_eq, y2 = timed(lambda: mimoAdaptEqualizer(y1, pq))
    t_cpr, y3 = timed(lambda: cpr(y2, pc))
    # symbol error rate of the recovered constellation (sanity of the whole chain)
    c = grayMapping(16, "qam").astype(np.complex128); c /= np.sqrt(np.mean(np.abs(c) ** 2))
    tail = slice(nsym // 2, nsym - 1000)

    # CPU oracle on a bounded sample
    n_cpu = 1 << a.cpu_nsym_log2
    xc = x[:2 * n_cpu]
    t0 = time.perf_counter(); y1c = ro.edc(xc, 800, 16, 193.1e12, 64e9, 32e9); tc_edc = time.perf_counter() - t0
    t0 = time.perf_counter()
    y2c, *_ = ro.mimo_adapt_equalizer(y1c, None, grayMapping(16, "qam"), nTaps=31, SpS=2, alg=["cma", "rde"],
                                      mu=[5e-3, 2e-4], L=[int(0.2 * n_cpu), int(0.8 * n_cpu)])
    tc_eq = time.perf_counter() - t0
    t0 = time.perf_counter(); ro.cpr_bps(y2c, grayMapping(16, "qam"), N=25, B=64, runFOE=False); tc_cpr = time.perf_counter() - t0
    ms = lambda n, t: n / t / 1e6
    # many independent streams (WDM channels x Monte-Carlo realisations): the regime the per-stream warp
    # design is for.  1184 streams (8 per SM) x 2^13 symbols, device-resident timing of the kernels only.
    from opticommpy_b200.equalization import mimoAdaptEqualizerBatch
    nb, lb = 1184, 1 << 13
    xb = [x[i * 64:i * 64 + 2 * lb] for i in range(nb)]
    pb = Bag(nTaps=31, SpS=2, M=16, constType="qam", alg=["cma", "rde"], mu=[5e-3, 2e-4],
             L=[lb // 4, lb - lb // 4], prgsBar=False)
    mimoAdaptEqualizerBatch(xb[:8], pb)
    torch.cuda.synchronize()
    from opticommpy_b200 import _cabi
    lib = _cabi.lib()
    inner, tk = lib.ocb_mimo_eq_run, []

    class Timed:  # kernel-only time of the stage launches (the Python batch shim is host-bound)
        def __call__(self, *a):
            torch.cuda.synchronize(); t0 = time.perf_counter(); r = inner(*a); torch.cuda.synchronize()
            tk.append(time.perf_counter() - t0)
            return r
    lib.ocb_mimo_eq_run = Timed()
    t0 = time.perf_counter(); yb = mimoAdaptEqualizerBatch(xb, pb); torch.cuda.synchronize(); t_batch = time.perf_counter() - t0
    lib.ocb_mimo_eq_run = inner
    t_kern = sum(tk)
    print(json.dumps({
        "workload": f"cfg3: edc(800 km, 448 taps) + 2x2 mimoAdaptEqualizer(CMA->RDE, 31 taps) + cpr/bps(B=64, N=25) on 2^{a.nsym_log2 + 1} samples x 2 pol",
        "gpu_input_Msamples_per_s": {"edc": ms(2 * nsym, t_edc), "mimoAdaptEqualizer": ms(2 * nsym, t_eq), "cpr_bps": ms(2 * nsym, t_cpr),
                                     "chain": ms(2 * nsym, t_edc + t_eq + t_cpr)},
        "gpu_seconds": {"edc": t_edc, "mimoAdaptEqualizer": t_eq, "cpr_bps": t_cpr},
        "cpu_oracle_input_Msamples_per_s": {"edc": ms(2 * n_cpu, tc_edc), "mimoAdaptEqualizer": ms(2 * n_cpu, tc_eq),
                                            "cpr_bps": ms(2 * n_cpu, tc_cpr), "chain": ms(2 * n_cpu, tc_edc + tc_eq + tc_cpr),
                                            "sample": f"2^{a.cpu_nsym_log2} symbols, 1 core"},
        "equalizer_batch": {"streams": nb, "symbols_per_stream": lb, "seconds_incl_host": t_batch,
                            "seconds_kernels_only": t_kern, "aggregate_Msym_per_s_kernels": nb * lb / t_kern / 1e6,
                            "aggregate_Msym_per_s": nb * lb / t_batch / 1e6,
                            "aggregate_input_Msamples_per_s": 2 * nb * lb / t_batch / 1e6},
        "api": "public drop-in calls, numpy in / numpy out (H2D + D2H inside)",
    }))


if __name__ == "__main__":
    main()
