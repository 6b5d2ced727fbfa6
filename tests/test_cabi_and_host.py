"""CPU-only checks of the boundary: the C-ABI library loads and exports every symbol the public
header declares; host-side logic (parameter bag, constellations, argument parsing) matches the
reference; the product path refuses to run without a GPU instead of falling back."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import Bag, rel_l2


def test_library_exports_every_declared_symbol():
    from opticommpy_b200 import _cabi
    lib = _cabi.lib()
    declared = _cabi.declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), name
    assert set(declared) == set(_cabi.SIGNATURES)
    assert lib.ocb_abi_version() == 1


def test_struct_layouts_match_the_header():
    from opticommpy_b200 import _cabi
    assert C.sizeof(_cabi.ManakovParams) == 8 * 8 + 6 * 4 + 3 * 8 + 2 * 4
    assert C.sizeof(_cabi.ManakovStats) == 5 * 8
    assert C.sizeof(_cabi.NlseParams) == 5 * 8 + 4 * 4 + 3 * 8


def test_argument_validation_without_gpu():
    from opticommpy_b200 import _cabi
    lib = _cabi.lib()
    assert lib.ocb_edc_workspace_bytes(1 << 16, 2, 448) > (1 << 16) * 2 * 8
    assert lib.ocb_edc_workspace_bytes(0, 2, 448) == -1
    rc = lib.ocb_mimo_eq_run(None, None, None, None, None, None, None, 1, 10, 0, 0, 0, 0, 0, 1, 2, 15, 2, 0, 0.0,
                             None, 0, None, 0, 0.0, 0, None)
    assert rc != 0 and b"NULL" in lib.ocb_last_error()


@pytest.mark.skipif(__import__("torch").cuda.is_available(), reason="only meaningful on a CPU-only box")
def test_no_cpu_fallback():
    from opticommpy_b200 import _cabi
    from opticommpy_b200.carrierRecovery import bps
    from opticommpy_b200.channels import manakovSSF
    x = np.ones((64, 2), dtype=complex)
    with pytest.raises(_cabi.OcbError):
        manakovSSF(x, Bag(Fs=64e9, prgsBar=False))
    with pytest.raises(_cabi.OcbError):
        bps(x, 2, np.array([1, -1], dtype=complex), 4)


def test_constellations_match_reference(golden):
    from opticommpy_b200.modulation import grayMapping
    for M in (4, 16, 64, 256):
        assert np.array_equal(grayMapping(M, "qam"), golden[f"const_qam{M}"])
    for M in (4, 8, 16):
        assert np.array_equal(grayMapping(M, "psk"), golden[f"const_psk{M}"])
    assert np.array_equal(grayMapping(16, "apsk"), golden["const_apsk16"])


def test_parameters_bag():
    from opticommpy_b200.utils import parameters
    p = parameters()
    p.Fs, p.taps = 64e9, [1, 2, 3]
    q = p.copy()
    q.taps.append(4)
    assert p.taps == [1, 2, 3] and q.Fs == 64e9
    assert p.to_engineering_notation(64e9) == "64.0 G"
    assert p.to_engineering_notation(12) == 12
    p.view(); p.table(); p.latex_table()


def test_legacy_noise_stream_matches_reference(golden):
    from opticommpy_b200._engine import legacy_complex_noise
    assert np.array_equal(legacy_complex_noise((2, 64), 3.0e-7, 5), golden["noise_seed5"])


def test_equalizer_argument_parsing(golden):
    from opticommpy_b200.equalization import _parse_equalizer_args
    s = _parse_equalizer_args(golden["eq_in"].T, Bag(nTaps=15, SpS=2, M=16, alg=["cma", "rde"], mu=[5e-3, 2e-3],
                                                     L=[1000, 2000]), None)
    assert s.nModes == 2 and s.nPad == 6000 + 14 and s.totalNumSymb == 3000 and s.symbRef is None
    assert s.H[0, 7] == 1 and s.H[3, 7] == 1 and np.count_nonzero(s.H) == 2
    assert s.mu.dtype == np.float32 and len(s.Rrde) == 3
    assert s.Rcma == pytest.approx(1.32, abs=1e-6)
    s = _parse_equalizer_args(golden["eq_in"], Bag(alg=["rls"], mu=[1e-3]), golden["eq_ref"])
    assert s.symbRef is not None and s.lambdaRLS == 0.99      # 'rls' trains against the reference symbols
    _parse_equalizer_args(golden["eq_in"], Bag(alg=["rls"], mu=[1e-3], nTaps=35), None)  # two matrix rows per lane
    with pytest.raises(NotImplementedError):                  # at most two matrix rows per lane: nTaps <= 64
        _parse_equalizer_args(golden["eq_in"], Bag(alg=["rls"], mu=[1e-3], nTaps=65), None)
    with pytest.raises(NotImplementedError):                  # rlsUp has no widely-linear update
        _parse_equalizer_args(golden["eq_in"], Bag(alg=["dd-rls"], mu=[1e-3], runWL=True), None)


def test_frontend_argument_errors():
    """Errors of the Rx front-end mirrors that the reference raises before any arithmetic (no GPU needed)."""
    from opticommpy_b200.core import decimate, firFilter
    x = np.ones((4001, 2), dtype=complex)
    with pytest.raises(ValueError):   # numpy's reshape(-1, SpSin) error in the reference (core.py:475)
        decimate(x, Bag(SpSin=4, SpSout=2))
    with pytest.raises(ValueError):
        firFilter(np.ones(5000), x)   # filter longer than the signal: not supported on the GPU path
    with pytest.raises(ValueError):
        firFilter(np.ones(0), x)


def test_bench_host_helpers():
    """bench.py / bench_extras.py pieces that need no GPU: the static config shared by both arms, the WDM input generator
    (reference transmitter when baseline/_ref or /root/reference is importable, else the synthetic stand-in) and the
    provenance of the roofline traffic figure."""
    import bench
    import bench_extras as bx
    assert bench.static_config(10, 1) == bench.static_config(10, 1) and "workload" in bench.static_config(10, 8)
    v, src = bench.ncu_traffic("fused")
    assert v and v > 1e7 and src["file"].startswith("profiles/") and src["captured"]
    sig, symb, grid, pulse, source = bx.wdm_waveform(3, 8, 8, seed=1)
    assert sig.shape == (256 * 8, 2) and symb.shape == (256, 2, 3) and len(grid) == 3 and len(pulse) == 1024
    p_tot = np.mean(np.sum(np.abs(sig) ** 2, axis=1))
    assert 0.5 * 3 * 10 ** (-0.2) * 1e-3 < p_tot < 2 * 3 * 10 ** (-0.2) * 1e-3     # 3 channels at -2 dBm
