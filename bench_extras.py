"""Extra workloads reported in bench.py's JSON line next to the headline (BASELINE.json configs[0], [2], [3], [4]).
None of them changes the headline numbers; each is bounded to a few seconds of GPU time.

  cfg1     : single-channel 2-pol manakovSSF, 2^16 samples, 1 span, hz = 0.8 km (101 executed steps) — small-N regime
  rx_chain : edc + 2x2 mimoAdaptEqualizer(CMA -> RDE, 31 taps) + cpr/bps(B = 64) on 2^22 samples x 2 pol, device-resident
             (rxChain: one upload, one download), per-stage device times, roofline figures, CPU port on a bounded sample
  cfg4_dbp : 11-channel WDM field after the cfg2 link, per-channel front end + manakovDBP, channels sharded over the ranks
  cfg5_mc  : 64 ASE-noise seeds x 5-channel WDM SSFM + receiver + on-device error counting, seeds sharded over the ranks,
             three scalars per seed gathered with one all_gather
  cfg2_concurrent : independent cfg2 realisations in flight on one GPU ; wdm_tx : the device-side transmitter (simpleWDMTx)
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))


class Bag:
    def __init__(self, **kw):
        self.__dict__.update(kw)

    def copy(self):
        return Bag(**self.__dict__)


# ---- WDM transmitter input (host side, outside every timed region) -----------------------------------------------------
def wdm_waveform(n_ch, nsym_log2, sps, seed=123, power_dbm=-2.0, spacing=37.5e9, rs=32e9):
    """(sig (N, 2) complex128, symb (nSym, 2, nCh), freqGrid, pulse, source): the reference's own simpleWDMTx
    (optic/models/tx.py:42-228, imported from baseline/_ref or /root/reference) when it can be imported, else the product's
    device transmitter opticommpy_b200.tx.simpleWDMTx (the same symbols for the same seed, the field within 2e-6)."""
    import bench
    ref = bench.import_reference()
    nsym = 1 << nsym_log2
    if ref is not None:
        try:
            from optic.dsp.core import pulseShape
            from optic.models.tx import simpleWDMTx
            _, parameters, where = ref
            p = parameters()
            p.M, p.Rs, p.SpS, p.nBits, p.pulseType, p.nFilterTaps, p.pulseRollOff = 16, rs, sps, 4 * nsym, "rrc", 1024, 0.01
            p.powerPerChannel, p.nChannels, p.Fc, p.wdmGridSpacing, p.nPolModes, p.seed, p.prgsBar = power_dbm, n_ch, 193.1e12, spacing, 2, seed, False
            sig, symb, p = simpleWDMTx(p)
            q = parameters()
            q.pulseType, q.nFilterTaps, q.rollOff, q.SpS = "rrc", 1024, 0.01, sps
            pulse = pulseShape(q)
            return sig, symb, np.asarray(p.wdmFreqGrid), pulse / np.max(np.abs(pulse)), f"optic.models.tx.simpleWDMTx ({where})"
        except Exception as e:  # fall through to the stand-in
            src_err = f" (simpleWDMTx failed: {e})"
    else:
        src_err = ""
    # no reference on this box: the product's own transmitter (same symbols for the same seed, field within 2e-6)
    from opticommpy_b200.tx import pulseShape as pulse_b200, simpleWDMTx as tx_b200
    p = Bag(M=16, Rs=rs, SpS=sps, nBits=4 * nsym, pulseType="rrc", nFilterTaps=1024, pulseRollOff=0.01, powerPerChannel=power_dbm,
            nChannels=n_ch, Fc=193.1e12, wdmGridSpacing=spacing, nPolModes=2, seed=seed, prgsBar=False)
    sig, symb, p = tx_b200(p)
    pulse = pulse_b200(Bag(pulseType="rrc", nFilterTaps=1024, rollOff=0.01, SpS=sps))
    return sig, symb, np.asarray(p.wdmFreqGrid), pulse / np.max(np.abs(pulse)), "opticommpy_b200.tx.simpleWDMTx (device)" + src_err


# ---- transmitter -------------------------------------------------------------------------------------------------------------
def extra_wdm_tx(torch, n_ch=11, nsym_log2=16, sps=16, cpu_nsym_log2=12):
    """The cfg2 / cfg4 input generator: 11-channel DP-16QAM, 2^16 symbols x 16 SpS = 2^20 samples per polarisation, through
    opticommpy_b200.tx.wdm_tx_rows_device (output stays on the GPU), next to the unmodified reference's simpleWDMTx on a
    bounded sample (one core) with the agreement of the two fields on that sample."""
    import bench
    from opticommpy_b200.tx import simpleWDMTx, wdm_tx_rows_device
    kw = dict(M=16, Rs=32e9, SpS=sps, pulseType="rrc", nFilterTaps=1024, pulseRollOff=0.01, powerPerChannel=-2.0, nChannels=n_ch,
              Fc=193.1e12, wdmGridSpacing=37.5e9, nPolModes=2, seed=123, prgsBar=False)
    nsym = 1 << nsym_log2
    wdm_tx_rows_device(Bag(nBits=4 * nsym, **kw))  # warm-up (cuFFT plans, allocator)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rows, symb, _ = wdm_tx_rows_device(Bag(nBits=4 * nsym, **kw))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    units = n_ch * 2 * nsym * sps  # (channel, mode) samples shaped and modulated
    res = {"workload": f"simpleWDMTx: {n_ch}-ch DP-16QAM, 2^{nsym_log2} symbols x {sps} SpS per (channel, mode), rrc 1024 taps",
           "seconds": dt, "value": units / dt / 1e6, "unit": "M (channel, mode) samples/s, host symbol draw + upload included",
           "output": [int(v) for v in rows.shape]}
    ref = bench.import_reference()
    if ref is not None:
        try:
            from optic.models.tx import simpleWDMTx as tx_ref
            _, parameters, where = ref
            n_cpu = 1 << cpu_nsym_log2
            p = parameters()
            for k, v in kw.items():
                setattr(p, k, v)
            p.nBits = 4 * n_cpu
            t0 = time.perf_counter()
            sig_ref, symb_ref, _ = tx_ref(p)
            t_ref = time.perf_counter() - t0
            sig_dev, symb_dev, _ = simpleWDMTx(Bag(nBits=4 * n_cpu, **kw))
            res["cpu_baseline"] = {"kind": "reference", "cores": 1, "sample": f"2^{cpu_nsym_log2} symbols per (channel, mode), {where}",
                                   "value": n_ch * 2 * n_cpu * sps / t_ref / 1e6, "seconds": t_ref,
                                   "rel_l2_device_vs_reference": float(np.linalg.norm(sig_dev - sig_ref) / np.linalg.norm(sig_ref)),
                                   "symbols_identical": bool(np.array_equal(symb_dev, symb_ref))}
        except Exception as e:
            res["cpu_baseline"] = {"error": f"{type(e).__name__}: {e}"}
    return res


def _sync(torch, dist, world):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def _max_over_ranks(torch, dist, world, ms):
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ---- cfg1 ----------------------------------------------------------------------------------------------------------------
def extra_cfg1(torch):
    import bench
    from opticommpy_b200.channels import manakov_rows_device
    n = 1 << 16
    rng = np.random.default_rng(8)
    x = (rng.normal(size=(n, 2)) + 1j * rng.normal(size=(n, 2))) * np.sqrt(11 * 10 ** (-0.2) * 1e-3 / 4)
    rows0 = torch.from_numpy(np.ascontiguousarray(x.T.astype(np.complex64))).cuda()
    prm = bench.channel_param(1, Fs=64e9, hz=0.8, amp=None)
    rows = rows0.clone()
    for _ in range(3):
        rows.copy_(rows0)
        st = manakov_rows_device(rows, prm, +1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps, ms = 10, 0.0
    for _ in range(reps):
        rows.copy_(rows0)
        e0.record()
        st = manakov_rows_device(rows, prm, +1)
        e1.record()
        torch.cuda.synchronize()
        ms += e0.elapsed_time(e1)
    res = {"workload": "cfg1: single-channel 2-pol manakovSSF, 2^16 samples, 1 span of 80 km, hz = 0.8 km, fixed step",
           "value": n * st["steps"] * reps / (ms * 1e-3) / 1e6, "unit": "Msamples/s", "steps": st["steps"],
           "iterations": st["iterations"], "us_per_ssfm_step": 1e3 * ms / reps / st["steps"]}
    # the same propagation for 8 independent waveforms in flight on the GPU (sharding.run_concurrent): at this size one
    # waveform is a latency chain through 64 small CTAs per launch, the chip's throughput shows with several at once
    from opticommpy_b200.sharding import run_concurrent

    def unit(i):
        r = rows0.clone()
        return manakov_rows_device(r, prm, +1)["steps"]

    run_concurrent(unit, list(range(8)), 8)
    torch.cuda.synchronize()
    e0.record()
    done = run_concurrent(unit, list(range(32)), 8)
    e1.record()
    torch.cuda.synchronize()
    res["eight_in_flight"] = {"value": n * sum(done.values()) / (e0.elapsed_time(e1) * 1e-3) / 1e6, "unit": "Msamples/s (aggregate)",
                              "waveforms": 32, "in_flight": 8}
    return res


# ---- cfg2 waveforms in flight ---------------------------------------------------------------------------------------------
def extra_cfg2_concurrent(torch, rows0, in_flight=(2, 3), spans=1):
    """Several independent realisations of the cfg2 waveform (N = 2^20, one span = 1001 steps each) propagated at the same
    time on ONE GPU through sharding.run_concurrent: one launch of the step loop is a single wave over the SMs whose
    load / transform / store phases do not overlap, so independent waveforms on separate streams fill each other's gaps.
    This is the Monte-Carlo / parameter-sweep regime, not the headline (which is one waveform per GPU)."""
    import bench
    from opticommpy_b200.channels import manakov_rows_device
    from opticommpy_b200.sharding import run_concurrent
    n = int(rows0.shape[1])
    prm = bench.channel_param(spans)

    def unit(i):
        r = rows0.clone()
        p = prm.copy() if hasattr(prm, "copy") else prm
        return manakov_rows_device(r, p, +1)["steps"]

    out = {"workload": f"cfg2 waveform, {spans} span(s) per realisation, R independent realisations in flight on one GPU "
                       "(one host thread, stream and plan each)", "unit": "Msamples/s (aggregate)"}
    for R in (1,) + tuple(in_flight):
        run_concurrent(unit, list(range(R)), R)  # warm-up: plans and tables of the worker streams
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        done = run_concurrent(unit, list(range(2 * R)), R)
        e1.record()
        torch.cuda.synchronize()
        out[f"in_flight_{R}"] = n * sum(done.values()) / (e0.elapsed_time(e1) * 1e-3) / 1e6
    return out


# ---- cfg3: receiver chain --------------------------------------------------------------------------------------------------
def extra_rx_chain(torch, peak_gbs, nsym_log2=21, cpu_nsym_log2=15):
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from cfg3_signal import make_signal
    from opticommpy_b200.modulation import grayMapping
    from opticommpy_b200.rxchain import rxChain
    c0 = grayMapping(16, "qam").astype(np.complex128)
    c = c0 / np.sqrt(np.mean(np.abs(c0) ** 2))
    nsym = 1 << nsym_log2
    x, _ = make_signal(nsym, c, seed=0)
    x = x.astype(np.complex64)
    pe = Bag(L=800, D=16, Fc=193.1e12, Fs=64e9, Rs=32e9)
    pq = lambda n: Bag(nTaps=31, SpS=2, M=16, constType="qam", alg=["cma", "rde"], mu=[1e-3, 2e-4],
                       L=[int(0.2 * n), int(0.8 * n)], prgsBar=False)
    pc = Bag(alg="bps", M=16, constType="qam", N=25, B=64, runFOE=False)
    rxChain(x, pe, pq(nsym), pc)  # warm-up at full size: cuFFT plan cache, caching allocator
    torch.cuda.synchronize()
    timing = {}
    t0 = time.perf_counter()
    out = rxChain(x, pe, pq(nsym), pc, timing=timing)
    t_e2e = time.perf_counter() - t0
    n_samp = 2 * nsym  # 2-pol input samples
    d = np.min(np.abs(out[nsym // 2:nsym - 1000, :, None] - c), axis=-1)
    stage = {k: timing[k] for k in ("h2d_pack", "edc", "equalizer", "cpr_bps")}
    # rooflines: edc moves 16 B per sample-mode algorithmically (complex64 read + write); bps evaluates B*M distances per
    # symbol-mode in float64 (~8 flop each: rotate 6 + 2 per point... counted as 8*B*M), against the FP64 vector peak
    edc_gbs = 16.0 * n_samp * 2 / (stage["edc"] * 1e-3) / 1e9
    bps_flops = 8.0 * 64 * 16 * nsym * 2
    res = {
        "workload": f"cfg3: edc(800 km, 448 taps) + 2x2 mimoAdaptEqualizer(CMA->RDE, 31 taps) + cpr/bps(B=64, N=25), 2^{nsym_log2 + 1} samples x 2 pol, one stream",
        "api": "opticommpy_b200.rxchain.rxChain: one H2D (complex64), stages device-resident, one D2H (complex128)",
        "chain_e2e_Msamples_per_s": n_samp / t_e2e / 1e6, "chain_e2e_seconds": t_e2e,
        "stage_device_ms": stage,
        "stage_Msamples_per_s": {k: n_samp / (v * 1e-3) / 1e6 for k, v in stage.items()},
        "h2d_bytes": timing["h2d_bytes"], "d2h_bytes": timing["d2h_bytes"],
        "edc_roofline": {"bound": "hbm", "bytes_per_sample_mode": 16, "achieved_GBps": edc_gbs, "peak_GBps": peak_gbs,
                         "frac": edc_gbs / peak_gbs,
                         "note": "cuFFT overlap-save: gather + batched FFT + multiply + IFFT + scatter = 5 passes over the blocks"},
        "bps_roofline": {"bound": "fp64", "flops": bps_flops, "achieved_TFLOPs": bps_flops / (stage["cpr_bps"] * 1e-3) / 1e12,
                         "peak_TFLOPs_nominal": 37.0, "note": "cpr stage time includes cast, unwrap scan and two pnorm passes"},
        "equalizer_note": "one 2x2x31 stream is a serial tap recurrence (latency-bound, one warp per output mode): symbols/s = clock / cycles per symbol",
        "equalizer_cycles_per_symbol_at_1965MHz": stage["equalizer"] * 1e-3 * 1.965e9 / nsym,
        "recovered_rms_distance_to_constellation": float(np.sqrt(np.mean(d ** 2))),
    }
    # many independent streams: the regime the per-stream warps are built for (same total number of samples)
    from opticommpy_b200.equalization import _parse_equalizer_args, _to_device, equalizer_stages_device
    nS, lb = 64, 1 << (nsym_log2 - 6)
    s0 = _parse_equalizer_args(np.zeros((2 * lb, 2), dtype=np.complex64), pq(lb), None)
    d_x = torch.zeros((nS, s0.nPad, 2, 2), dtype=torch.float32, device="cuda")
    xs = torch.from_numpy(np.ascontiguousarray(x[: nS * 2 * lb].reshape(nS, 2 * lb, 2)).view(np.float32).reshape(nS, 2 * lb, 2, 2)).cuda()
    d_x[:, s0.Lpad:s0.Lpad + 2 * lb] = xs
    d_H = _to_device(torch, np.stack([s0.H] * nS).view(np.float32))
    equalizer_stages_device(s0, nS, d_x, None, 0, d_H.clone(), None)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    equalizer_stages_device(s0, nS, d_x, None, 0, d_H, None)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    res["equalizer_64_streams"] = {"streams": nS, "symbols_per_stream": lb, "device_ms": ms,
                                   "aggregate_input_Msamples_per_s": nS * 2 * lb / (ms * 1e-3) / 1e6}
    # CPU port (oracle) on a bounded sample, one core
    try:
        from oracle import rxdsp_oracle as ro
        n_cpu = 1 << cpu_nsym_log2
        xc = x[: 2 * n_cpu].astype(np.complex128)
        t0 = time.perf_counter(); y1 = ro.edc(xc, 800, 16, 193.1e12, 64e9, 32e9); t_a = time.perf_counter() - t0
        t0 = time.perf_counter()
        y2, *_ = ro.mimo_adapt_equalizer(y1, None, c0, nTaps=31, SpS=2, alg=["cma", "rde"], mu=[1e-3, 2e-4], L=[int(0.2 * n_cpu), int(0.8 * n_cpu)])
        t_b = time.perf_counter() - t0
        t0 = time.perf_counter(); ro.cpr_bps(y2, c0, N=25, B=64, runFOE=False); t_c = time.perf_counter() - t0
        res["cpu_baseline"] = {"kind": "port", "cores": 1, "sample": f"2^{cpu_nsym_log2} symbols x 2 pol (oracle: numpy edc, plain-C equalizer, numpy/C bps)",
                               "chain_Msamples_per_s": 2 * n_cpu / (t_a + t_b + t_c) / 1e6,
                               "stage_Msamples_per_s": {"edc": 2 * n_cpu / t_a / 1e6, "equalizer": 2 * n_cpu / t_b / 1e6, "cpr_bps": 2 * n_cpu / t_c / 1e6}}
    except Exception as e:
        res["cpu_baseline"] = {"error": str(e)}
    return res


# ---- cfg4: per-channel DBP, channels sharded --------------------------------------------------------------------------------
def unit_workers(n_units):
    """Units (channels, seeds) kept in flight per GPU by sharding.run_concurrent in the sharded extras: the balanced
    default (at most 8), or OCB_UNIT_WORKERS."""
    from opticommpy_b200.sharding import balanced_workers
    v = os.environ.get("OCB_UNIT_WORKERS")
    return max(1, int(v)) if v else balanced_workers(max(1, n_units))


def extra_cfg4_dbp(torch, dist, world, rank, spans=10, hz_dbp=0.8):
    import bench
    from opticommpy_b200.channels import manakov_rows_device
    from opticommpy_b200.pipelines import dbp_channel_device, dbp_channels_device, upload_field
    from opticommpy_b200.sharding import gather_device, shard_units
    n_ch, sps, rs = 11, 16, 32e9
    fs = rs * sps
    sig, symb, grid, pulse, source = wdm_waveform(n_ch, 16, sps, seed=123)
    rows = upload_field(sig)
    fwd = bench.channel_param(spans)  # the cfg2 link (EDFA, seed 456) — identical on every rank
    manakov_rows_device(rows, fwd, +1)
    prm = Bag(Fs=2 * rs, Ltotal=80 * spans, Lspan=80, hz=hz_dbp, alpha=0.2, D=16, gamma=1.3, Fc=193.1e12, amp="edfa", NF=4.5,
              maxIter=10, tol=1e-5, nlprMethod=False, maxNlinPhaseRot=2e-2, seed=None)
    mine = shard_units(n_ch, rank, world)
    nw = unit_workers(len(mine))
    run = lambda units: dbp_channels_device(rows, grid, fs, pulse, sps, prm, units=units, workers=nw)
    run([n_ch // 2] * nw)  # warm-up (plans and tables of every worker stream)
    _sync(torch, dist, world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    done = run(mine)  # up to nw channels in flight on this GPU, one stream + plan each
    local = {k: v[0] for k, v in done.items()}
    steps = sum(v[1]["steps"] for v in done.values())
    full = gather_device(local, n_ch)  # the path's only collective: 11 x (2, 2^17) complex64 fields over NVLink
    e1.record()
    _sync(torch, dist, world)
    my_ms = e0.elapsed_time(e1)
    ms = _max_over_ranks(torch, dist, world, my_ms)
    tot_steps = torch.tensor([float(steps)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tot_steps)
    n2 = int(full[0].shape[1])
    sizes = [len(shard_units(n_ch, r, world)) for r in range(world)]
    return {"workload": f"cfg4: 11-ch WDM after the cfg2 link ({spans} x 80 km), per channel: CW-LO coherent front end (down-shift) + matched filter + "
                        f"decimate 16->2 SpS + manakovDBP ({spans} spans, hz = {hz_dbp} km, N = 2^17), one channel per shard unit",
            "input": source, "value": n2 * float(tot_steps.item()) / (ms * 1e-3) / 1e6, "unit": "Msamples/s (DBP sample-steps)",
            "seconds": ms * 1e-3, "channels": n_ch, "shard_sizes": sizes, "units_in_flight_per_gpu": nw,
            "balance_bound": n_ch / (max(sizes) * world), "dbp_steps_total": int(tot_steps.item()),
            "gathered": [int(v) for v in full[0].shape], "gather": "one all_gather of the device tensors (no host round trip)"}


# ---- cfg5: Monte-Carlo seeds sharded -------------------------------------------------------------------------------------------
def extra_cfg5_mc(torch, dist, world, rank, n_seeds=64, spans=3, hz=0.1):
    import bench
    from opticommpy_b200.channels import manakov_rows_device
    from opticommpy_b200.core import symbolSync
    from opticommpy_b200.pipelines import RxRecipe, channel_frontend_device, monte_carlo_ber_device, rx_symbols_device, upload_field
    from opticommpy_b200.equalization import edc_rows_device
    from opticommpy_b200.sharding import gather_device, shard_units
    n_ch, sps, rs = 5, 8, 32e9
    fs = rs * sps
    sig, symb, grid, pulse, source = wdm_waveform(n_ch, 15, sps, seed=321)
    n = len(sig)
    rows0 = upload_field(sig)
    base = Bag(Fs=fs, Ltotal=80 * spans, Lspan=80, hz=hz, alpha=0.2, D=16, gamma=1.3, Fc=193.1e12, amp="edfa", NF=4.5,
               maxIter=10, tol=1e-5, nlprMethod=False, maxNlinPhaseRot=2e-2, seed=1000)
    ch = n_ch // 2
    # seed-independent alignment of the reference symbols (symbolSync on the first realisation's 2-SpS signal)
    rows = rows0.clone()
    manakov_rows_device(rows, base, +1)
    rec = RxRecipe(fs, rs, sps, pulse, 80 * spans, 16, 193.1e12, symb[:, :, ch], mu=(2e-2, 2e-3))
    s2 = channel_frontend_device(rows, float(grid[ch]), fs, pulse, sps, 2)
    d_in = torch.view_as_real(s2).contiguous()
    d_edc = torch.empty_like(d_in)
    _k = edc_rows_device(d_in, d_edc, rec.h_edc)
    s2h = torch.view_as_complex(d_edc).cpu().numpy().T
    txs = symbolSync(s2h, symb[:, :, ch], 2, "amp")
    rec.symbRef = np.ascontiguousarray((txs / np.sqrt(np.mean(np.abs(txs) ** 2))).astype(np.complex64))
    rx_symbols_device(rows, float(grid[ch]), rec)  # warm-up
    mine = shard_units(n_seeds, rank, world)
    nw = unit_workers(len(mine))
    warm = base.copy()
    warm.Ltotal = 80
    monte_carlo_ber_device(rows0, [1000 + i for i in range(nw)], warm, float(grid[ch]), rec, workers=nw)  # warm-up of every worker stream
    _sync(torch, dist, world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    done = monte_carlo_ber_device(rows0, [1000 + i for i in mine], base, float(grid[ch]), rec, workers=nw)  # up to nw seeds in flight
    local = {i: done[1000 + i][0] for i in mine}
    steps = sum(v[1] for v in done.values())
    full = gather_device(local, n_seeds)  # 3 scalars per seed
    e1.record()
    _sync(torch, dist, world)
    ms = _max_over_ranks(torch, dist, world, e0.elapsed_time(e1))
    tot_steps = torch.tensor([float(steps)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tot_steps)
    res = torch.stack(full).cpu().numpy()
    return {"workload": f"cfg5: {n_seeds} ASE-noise seeds x 5-ch WDM DP-16QAM (N = 2^18, {spans} x 80 km, hz = {hz} km) manakovSSF + centre-channel "
                        "receiver (front end, matched filter, decimate, edc, nlms->dd-lms equalizer, bps) + on-device BER/SER/SNR",
            "input": source, "value": n * float(tot_steps.item()) / (ms * 1e-3) / 1e6, "unit": "Msamples/s (SSFM sample-steps, receiver time included)",
            "seconds": ms * 1e-3, "seeds": n_seeds, "seeds_per_s": n_seeds / (ms * 1e-3), "units_in_flight_per_gpu": nw,
            "shard_sizes": [len(shard_units(n_seeds, r, world)) for r in range(world)],
            "gathered": [int(v) for v in res.shape], "gather": "one all_gather of 3 float64 per seed",
            "ber_mean": float(res[:, 0].mean()), "ser_mean": float(res[:, 1].mean()),
            "snr_db_mean": float(res[:, 2].mean()), "snr_db_std": float(res[:, 2].std())}
