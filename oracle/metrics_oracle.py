"""ORACLE (test infrastructure, NOT product code) — CPU restatement in numpy of the reference's hard
decisions and Monte-Carlo error counting: ``minEuclid`` (optic/comm/modulation.py:271-299),
``demodulateGray`` + ``demap`` (:369-408, :303-333) and ``fastBERcalc`` (optic/comm/metrics.py:110-195).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this module.
Pinned against the reference through ``tests/golden/ref_metrics.npz`` (made by
``tests/golden/make_golden_metrics.py``).  The constellation is an argument (its ordering is pinned
separately by the ``const_*`` golden vectors), so this file depends on numpy alone.
"""
from __future__ import annotations

import numpy as np


def min_euclid(symb, const, chunk=1 << 14):
    """Index of the constellation point closest to each symbol; first index on ties (modulation.py:296-298)."""
    symb = np.asarray(symb).reshape(-1)
    const = np.asarray(const).reshape(-1)
    idx = np.empty(symb.shape[0], dtype=np.int64)
    for s in range(0, symb.shape[0], chunk):
        d = np.abs(symb[s:s + chunk, None] - const[None, :])
        idx[s:s + chunk] = np.argmin(d, axis=1)
    return idx


def index_bits(idx, bits_per_symbol):
    """Bits of every index, most significant first, flattened (modulation.py:396-407: the bit map of the
    Gray-ordered constellation is the binary expansion of the index, because ``minEuclid(const, const)`` is the
    identity)."""
    shifts = np.arange(bits_per_symbol - 1, -1, -1)
    return ((np.asarray(idx)[:, None] >> shifts[None, :]) & 1).reshape(-1).astype(np.int64)


def demodulate_gray(symb, const):
    b = int(np.log2(len(const)))
    return index_bits(min_euclid(symb, const), b)


def _columns(x):
    x = np.array(x)  # copy (metrics.py:154-155)
    if x.ndim == 1:
        return x.reshape(-1, 1)
    return x.T.copy() if x.shape[1] > x.shape[0] else x


def fast_ber_calc(rx, tx, const, constType, px=None, return_counts=False):
    """BER, SER and SNR estimate per column (metrics.py:110-195)."""
    const = np.asarray(const)
    M = len(const)
    b = int(np.log2(M))
    if px is None or len(px) == 0:
        px = np.ones(M) / M
    Es = np.sum(np.abs(const) ** 2 * px)  # :151
    rx, tx = _columns(rx), _columns(tx)
    nModes = tx.shape[1]
    BER, SER, SNR = np.zeros(nModes), np.zeros(nModes), np.zeros(nModes)
    counts = np.zeros((2, nModes), dtype=np.int64)
    for k in range(nModes):
        r, t = rx[:, k], tx[:, k]
        if constType in ("qam", "psk"):
            r = np.mean(t / r) * r  # phase-ambiguity correction (:176-179)
        r = r / np.sqrt(np.mean(r * np.conj(r)).real)  # pnorm (:181-182)
        t = t / np.sqrt(np.mean(t * np.conj(t)).real)
        with np.errstate(divide="ignore"):
            SNR[k] = 10 * np.log10(np.mean(np.abs(t) ** 2) / np.mean(np.abs(r - t) ** 2))  # :185
        a = min_euclid(np.sqrt(Es) * r, const)
        c = min_euclid(np.sqrt(Es) * t, const)
        diff = a ^ c
        counts[0, k] = sum(int(np.count_nonzero((diff >> s) & 1)) for s in range(b))
        counts[1, k] = int(np.count_nonzero(diff))
        BER[k] = counts[0, k] / (len(a) * b)  # :190-192
        SER[k] = counts[1, k] / len(a)
    if return_counts:
        return BER, SER, SNR, counts
    return BER, SER, SNR
