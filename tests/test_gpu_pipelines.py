"""Device-resident pipelines of BASELINE configs[3] / [4] (opticommpy_b200.pipelines) against the same chain built from
the CPU oracle's restatements of the reference functions: one upload, every stage on the GPU, one download."""
import os
import sys

import numpy as np
import pytest

from conftest import Bag, rel_l2

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _wdm(n_ch, nsym_log2, sps, seed):
    sys.path.insert(0, ROOT)
    import bench_extras as bx
    return bx.wdm_waveform(n_ch, nsym_log2, sps, seed=seed)


def test_cfg4_pipeline_ssf_frontend_dbp_vs_oracle():
    """SSF -> coherent front end (CW LO on the channel) -> matched filter -> decimate -> manakovDBP, device-resident."""
    import torch
    from opticommpy_b200.channels import manakov_rows_device
    from opticommpy_b200.pipelines import dbp_channel_device, upload_field
    from oracle import fiber_oracle as fo
    from oracle import frontend_oracle as fe
    n_ch, sps, rs = 3, 16, 32e9
    fs = rs * sps
    sig, symb, grid, pulse, _ = _wdm(n_ch, 12, sps, 11)          # N = 2^16
    n = len(sig)
    fwd = Bag(Fs=fs, Ltotal=160, Lspan=80, hz=2.0, alpha=0.2, D=16, gamma=1.3, Fc=193.1e12, amp="ideal", NF=4.5, maxIter=10,
              tol=1e-5, nlprMethod=False, maxNlinPhaseRot=2e-2, seed=None)
    rows = upload_field(sig)                                       # the one H2D copy
    manakov_rows_device(rows, fwd, +1)
    dbp = Bag(Fs=2 * rs, Ltotal=160, Lspan=80, hz=4.0, alpha=0.2, D=16, gamma=1.3, Fc=193.1e12, amp="ideal", NF=4.5,
              maxIter=10, tol=1e-5, nlprMethod=False, maxNlinPhaseRot=2e-2, seed=None)
    k = 2
    out, st = dbp_channel_device(rows, float(grid[k]), fs, pulse, sps, dbp)
    got = out.cpu().numpy().T                                      # the one D2H copy
    y = fo.manakov(sig, fo.FiberConfig(Fs=fs, Ltotal=160, Lspan=80, hz=2.0, amp="ideal", nlprMethod=False))
    lo = np.sqrt(1e-2) * np.exp(2j * np.pi * grid[k] * np.arange(n) / fs)
    s = fe.pdm_coherent_receiver_ideal(y, lo, fs)
    s = fe.fir_filter(pulse, s)
    s2, _ = fe.decimate(s, sps, 2)
    st_o = {}
    ref = fo.manakov(s2, fo.FiberConfig(Fs=2 * rs, Ltotal=160, Lspan=80, hz=4.0, amp="ideal", nlprMethod=False), direction=-1, stats=st_o)
    assert got.shape == ref.shape
    assert st["steps"] == st_o["steps"] and st["iterations"] == st_o["iterations"]
    assert rel_l2(got, ref) < 2e-4   # complex64 through five stages (front end at 1e-6, 1024-tap filter, 80 DBP steps)


def test_cfg5_receiver_and_error_counting_vs_oracle():
    """One Monte-Carlo unit: manakovSSF with the reference's seeded ASE realisation -> centre-channel receiver -> BER / SER /
    SNR counted on the device, against the oracle chain."""
    import torch
    from opticommpy_b200 import _engine
    from opticommpy_b200.channels import _edfa_numbers, manakov_rows_device
    from opticommpy_b200.core import symbolSync
    from opticommpy_b200.modulation import grayMapping
    from opticommpy_b200.pipelines import RxRecipe, ber_scalars_device, rx_symbols_device, upload_field
    from oracle import fiber_oracle as fo
    from oracle import frontend_oracle as fe
    from oracle import metrics_oracle as mo
    from oracle import rxdsp_oracle as ro
    n_ch, sps, rs = 5, 8, 32e9
    fs = rs * sps
    sig, symb, grid, pulse, _ = _wdm(n_ch, 13, sps, 5)           # N = 2^16, 8192 symbols
    n = len(sig)
    ch = n_ch // 2
    cfg = fo.FiberConfig(Fs=fs, Ltotal=160, Lspan=80, hz=2.0, amp="edfa", seed=7, nlprMethod=False)
    y = fo.manakov(sig, cfg)
    # oracle receiver
    lo = np.sqrt(1e-2) * np.exp(2j * np.pi * grid[ch] * np.arange(n) / fs)
    s = fe.fir_filter(pulse, fe.pdm_coherent_receiver_ideal(y, lo, fs))
    s2, _ = fe.decimate(s, sps, 2)
    s2 = ro.edc(s2, 160, 16, 193.1e12, 2 * rs, rs)
    s2 = s2 / np.sqrt(np.mean(np.abs(s2) ** 2))
    txs = fe.symbol_sync(s2, symb[:, :, ch], 2, "amp")
    txs = txs / np.sqrt(np.mean(np.abs(txs) ** 2))
    c0 = grayMapping(16, "qam")
    nsym = len(s2) // 2
    ntr = int(0.2 * nsym)
    yq, *_ = ro.mimo_adapt_equalizer(s2, txs, c0, nTaps=15, SpS=2, alg=["nlms", "dd-lms"], mu=[2e-2, 2e-3], L=[ntr, nsym - ntr])
    y3 = ro.cpr_bps(yq, c0, N=25, B=64, runFOE=False)[0]
    L = min(len(y3), len(txs))
    ber_o, ser_o, snr_o = mo.fast_ber_calc(y3[2000:L - 2000], txs[2000:L - 2000], c0, "qam")
    # device chain with the same ASE realisation (the reference's seeded MT19937 stream, injected)
    prm = Bag(Fs=fs, Ltotal=160, Lspan=80, hz=2.0, alpha=0.2, D=16, gamma=1.3, Fc=193.1e12, amp="edfa", NF=4.5, maxIter=10,
              tol=1e-5, nlprMethod=False, maxNlinPhaseRot=2e-2, seed=7)
    _, nvar = _edfa_numbers(0.2 * 80, 4.5, 193.1e12, fs)
    noise = torch.from_numpy(_engine.legacy_complex_noise((1, n), nvar, 7).astype(np.complex64)).cuda()
    rows = upload_field(sig)
    manakov_rows_device(rows, prm, +1, noise_rows=noise)
    assert rel_l2(rows.cpu().numpy().T, y) < 1e-4
    rec = RxRecipe(fs, rs, sps, pulse, 160, 16, 193.1e12, symb[:, :, ch], mu=(2e-2, 2e-3))
    # the aligned reference symbols come from the device symbolSync on the host copy of the oracle's 2-SpS signal
    txs_d = symbolSync(s2, symb[:, :, ch], 2, "amp")
    assert np.array_equal(txs_d, fe.symbol_sync(s2, symb[:, :, ch], 2, "amp"))
    rec.symbRef = np.ascontiguousarray(txs.astype(np.complex64))
    d_sym = rx_symbols_device(rows, float(grid[ch]), rec)
    sym = d_sym.cpu().numpy()
    assert rel_l2(sym[2000:L - 2000], y3[2000:L - 2000]) < 5e-3     # adaptive chain in complex64 vs float64
    ber, ser, snr = ber_scalars_device(d_sym, rec)
    assert abs(snr - float(np.mean(snr_o))) < 0.1
    assert abs(ber - float(np.mean(ber_o))) <= 2e-4 and ber < 1e-2
    # the sweep driver: several Philox-seeded units in flight == the same units one at a time (bit-identical scalars)
    from opticommpy_b200.pipelines import monte_carlo_ber_device
    rows0 = upload_field(sig)
    seeds = [11, 12, 13]
    con = monte_carlo_ber_device(rows0, seeds, prm, float(grid[ch]), rec, workers=3)
    seq = monte_carlo_ber_device(rows0, seeds, prm, float(grid[ch]), rec, workers=1)
    for sd in seeds:
        assert con[sd][1] == seq[sd][1] == 80
        assert torch.equal(con[sd][0], seq[sd][0])
        assert abs(float(con[sd][0][2]) - snr) < 0.5                 # another noise realisation of the same link
    assert not torch.equal(con[11][0], con[12][0])


def test_run_concurrent_units_equal_sequential_units():
    """sharding.run_concurrent: independent seeded propagations + front end + EDC with several units in flight on one GPU
    (one host thread, stream and plan per worker) give bit-identical results to the same units run one after the other."""
    import torch
    from opticommpy_b200.channels import manakov_rows_device
    from opticommpy_b200.equalization import _edc_taps, edc_rows_device
    from opticommpy_b200.pipelines import channel_frontend_device, upload_field
    from opticommpy_b200.sharding import run_concurrent
    n_ch, sps, rs = 3, 8, 32e9
    fs = rs * sps
    sig, symb, grid, pulse, _ = _wdm(n_ch, 13, sps, 5)           # N = 2^16
    rows0 = upload_field(sig)
    base = Bag(Fs=fs, Ltotal=160, Lspan=80, hz=1.0, alpha=0.2, D=16, gamma=1.3, Fc=193.1e12, amp="edfa", NF=4.5, maxIter=10,
               tol=1e-5, nlprMethod=False, maxNlinPhaseRot=2e-2, seed=None)
    h_edc, _, _ = _edc_taps(Bag(L=160, D=16, Fc=193.1e12, Rs=rs, Fs=2 * rs), 2 * rs)

    def unit(i):
        r = rows0.clone()
        p = Bag(**base.__dict__)
        p.seed = 77 + i                                           # on-device Philox noise keyed by the seed
        st = manakov_rows_device(r, p, +1)
        ch = channel_frontend_device(r, float(grid[1]), fs, pulse, sps, 2)
        d_in = torch.view_as_real(ch).contiguous()
        d_out = torch.empty_like(d_in)
        keep = edc_rows_device(d_in, d_out, h_edc)
        return d_out, st

    units = list(range(6))
    seq = {i: unit(i) for i in units}
    torch.cuda.synchronize()
    con = run_concurrent(unit, units, workers=3)
    torch.cuda.synchronize()
    assert sorted(con) == units
    for i in units:
        assert con[i][1] == seq[i][1]
        assert torch.equal(con[i][0], seq[i][0]), f"unit {i} differs between concurrent and sequential execution"
    assert not torch.equal(con[0][0], con[1][0])                  # different seeds, different noise
    # a failing unit surfaces in the caller
    def bad(i):
        if i == 2:
            raise ValueError("unit 2 failed")
        return unit(i)
    with pytest.raises(ValueError):
        run_concurrent(bad, units, workers=3)
