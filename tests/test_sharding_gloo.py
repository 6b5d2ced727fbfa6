"""Multi-process sharding logic on CPU: world_size 2 and 3 with the gloo backend (127.0.0.1).
The per-unit "work" is the CPU oracle here (tests may use it); on the GPU box the same driver
code runs the CUDA path over NCCL."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_units, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from opticommpy_b200 import sharding
    from oracle import fiber_oracle as fo

    rng = np.random.default_rng(0)
    units = [(rng.normal(size=(256, 2)) + 1j * rng.normal(size=(256, 2))) * 0.03 for _ in range(n_units)]
    def work(x, seed):  # a private config per call: units may run on concurrent host threads below
        return fo.manakov(x, fo.FiberConfig(Fs=64e9, Ltotal=20, Lspan=20, hz=5.0, amp="edfa", nlprMethod=False, seed=seed))

    mine = sharding.shard_units(n_units)
    local = {i: work(units[i], 100 + i) for i in mine}
    full = sharding.gather_results(local, n_units)
    ref = [work(units[i], 100 + i) for i in range(n_units)]
    ok = all(np.array_equal(a, b) for a, b in zip(full, ref)) and len(full) == n_units
    # the device-side gather (no numpy round trip; CPU tensors under gloo, CUDA tensors under NCCL): ragged / empty shards
    tl = {i: torch.from_numpy(local[i].astype(np.complex64)) for i in mine}
    tfull = sharding.gather_device(tl, n_units)
    ok = ok and len(tfull) == n_units and all(np.array_equal(t.numpy(), r.astype(np.complex64)) for t, r in zip(tfull, ref))
    # run_sharded with two of the rank's units in flight (host threads on this CPU box) gives the same gathered list
    full2 = sharding.run_sharded(lambda iu: work(units[iu], 100 + iu), list(range(n_units)), in_flight=2)
    ok = ok and len(full2) == n_units and all(np.array_equal(a, b) for a, b in zip(full2, ref))
    sc = {i: torch.tensor([float(i), 2.0 * i, -1.0], dtype=torch.float64) for i in mine}   # three scalars per unit (cfg5)
    sfull = sharding.gather_device(sc, n_units)
    ok = ok and all(s.tolist() == [float(i), 2.0 * i, -1.0] for i, s in enumerate(sfull))
    q.put((rank, mine, ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_units", [(2, 5), (3, 4), (2, 1)])
def test_shard_and_gather(world, n_units):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_units, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    owned = sorted(i for _, mine, _ in res for i in mine)
    assert owned == list(range(n_units))  # every unit exactly once
    assert all(ok for _, _, ok in res)    # every rank holds the complete, correct result list


def test_partition_properties():
    from opticommpy_b200.sharding import owner_of, shard_units
    assert [len(shard_units(11, r, 8)) for r in range(8)] == [2, 2, 2, 1, 1, 1, 1, 1]  # cfg4: 11 channels / 8 GPUs
    assert [len(shard_units(64, r, 8)) for r in range(8)] == [8] * 8                   # cfg5: 64 seeds / 8 GPUs
    for n, w in [(11, 8), (64, 8), (3, 5), (1, 2)]:
        for r in range(w):
            for i in shard_units(n, r, w):
                assert owner_of(i, n, w) == r


def _dbp_worker(rank, world, port, q):
    """cfg4 in miniature, end to end under gloo: a 3-channel WDM field, per-channel coherent front end + matched filter +
    decimation + digital back-propagation (the oracle's restatements stand in for the CUDA calls on this CPU box), channels
    sharded raggedly over the ranks (2 + 1), one gather of the back-propagated fields."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench_extras as bx
    from opticommpy_b200 import sharding
    from oracle import fiber_oracle as fo
    from oracle import frontend_oracle as fe

    n_ch, sps, rs = 3, 8, 32e9
    fs = rs * sps
    sig, symb, grid, pulse, _ = bx.wdm_waveform(n_ch, 9, sps, seed=3)      # identical on every rank (seeded)
    n = len(sig)
    y = fo.manakov(sig, fo.FiberConfig(Fs=fs, Ltotal=80, Lspan=80, hz=10.0, amp="ideal", nlprMethod=False))

    def unit(k):
        lo = np.sqrt(1e-2) * np.exp(2j * np.pi * grid[k] * np.arange(n) / fs)
        s = fe.fir_filter(pulse, fe.pdm_coherent_receiver_ideal(y, lo, fs))
        s2, _ = fe.decimate(s, sps, 2)
        return fo.manakov(s2, fo.FiberConfig(Fs=2 * rs, Ltotal=80, Lspan=80, hz=20.0, amp="ideal", nlprMethod=False), direction=-1)

    mine = sharding.shard_units(n_ch)
    local = {k: unit(k) for k in mine}
    full = sharding.gather_results(local, n_ch)
    tfull = sharding.gather_device({k: torch.from_numpy(v) for k, v in local.items()}, n_ch)
    ref = [unit(k) for k in range(n_ch)]
    ok = all(np.array_equal(a, b) for a, b in zip(full, ref)) and all(np.array_equal(t.numpy(), b) for t, b in zip(tfull, ref))
    q.put((rank, mine, ok))
    dist.barrier()
    dist.destroy_process_group()


def test_ragged_dbp_pipeline_end_to_end():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dbp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(len(m) for _, m, _ in res) == [1, 2]      # 3 channels over 2 ranks
    assert all(ok for _, _, ok in res)


def test_run_concurrent_scheduler_on_host_threads():
    """Host logic of sharding.run_concurrent (cuda=False): every unit runs exactly once, results are keyed by unit, at most
    `workers` units are in flight, and the first exception of a worker reaches the caller."""
    import threading
    import time

    from opticommpy_b200.sharding import run_concurrent
    lock, state = threading.Lock(), {"now": 0, "peak": 0, "calls": 0}

    def fn(u):
        with lock:
            state["now"] += 1
            state["calls"] += 1
            state["peak"] = max(state["peak"], state["now"])
        time.sleep(0.01)
        with lock:
            state["now"] -= 1
        return u * u

    out = run_concurrent(fn, range(13), workers=3, cuda=False)
    assert out == {u: u * u for u in range(13)}
    assert state["calls"] == 13 and 1 <= state["peak"] <= 3
    assert run_concurrent(fn, [], workers=3, cuda=False) == {}
    assert run_concurrent(fn, [5], workers=8, cuda=False) == {5: 25}

    def bad(u):
        if u == 4:
            raise KeyError("unit 4")
        return u

    import pytest
    with pytest.raises(KeyError):
        run_concurrent(bad, range(8), workers=2, cuda=False)


def test_balanced_workers():
    """Fewest workers (<= 8) that finish n equal units in the minimal number of rounds."""
    from opticommpy_b200.sharding import balanced_workers
    assert [balanced_workers(n) for n in (0, 1, 2, 8, 9, 11, 16, 17, 64)] == [1, 1, 2, 8, 5, 6, 8, 6, 8]
    for n in range(1, 200):
        w = balanced_workers(n)
        assert 1 <= w <= 8 and -(-n // w) == -(-n // 8)       # never more rounds than with 8 workers
        assert w == 1 or -(-n // (w - 1)) > -(-n // 8)          # and no smaller pool achieves that
    assert balanced_workers(11, max_workers=4) == 4 and balanced_workers(5, max_workers=2) == 2
