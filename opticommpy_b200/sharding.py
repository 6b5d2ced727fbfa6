"""Multi-GPU sharding of independent propagation / DSP units.

The hot path has no data-path collective: WDM channels for back-propagation, Monte-Carlo noise
seeds and power-sweep points are independent waveforms.  Units are block-partitioned over the
ranks of one ``torch.distributed`` job (one process per GPU, NCCL over NVLink on the GPU box, gloo
in the CPU tests), every rank runs its units with the single-GPU functions of this package, and
ONE collective at the end returns the results to every rank (``all_gather``) — the only
communication of the path.

    units = [...]                                  # identical list on every rank
    mine  = shard_units(len(units))                # indices this rank owns
    local = {i: manakovSSF(units[i], param.copy()) for i in mine}
    full  = gather_results(local, len(units))      # list of numpy arrays, same on every rank

Note on semantics: the reference's K-column batching couples the columns of one call through
``max(phiRot)`` and the Frobenius norm (channels.py:394, 517), so independent waveforms are
separate calls here (K = 1 each), which reproduces the per-waveform reference results.
"""
from __future__ import annotations

import numpy as np


def _dist():
    import torch.distributed as dist

    return dist if (dist.is_available() and dist.is_initialized()) else None


def world():
    d = _dist()
    return (d.get_rank(), d.get_world_size()) if d else (0, 1)


def shard_units(n_units: int, rank: int | None = None, world_size: int | None = None):
    """Contiguous block partition of ``range(n_units)``; the first ``n_units % world`` ranks get one
    extra unit (11 channels over 8 GPUs -> 2,2,2,1,1,1,1,1)."""
    if rank is None or world_size is None:
        rank, world_size = world()
    base, extra = divmod(n_units, world_size)
    start = rank * base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))


def owner_of(unit: int, n_units: int, world_size: int) -> int:
    base, extra = divmod(n_units, world_size)
    edge = extra * (base + 1)
    return unit // (base + 1) if unit < edge else extra + (unit - edge) // max(base, 1)


def gather_results(local: dict, n_units: int, device=None):
    """All-gather per-unit numpy results.  ``local`` maps unit index -> ndarray; every unit must have
    the same shape and dtype.  Ragged ownership (n_units not divisible by the world size) is padded
    to the largest shard so that a single equal-size ``all_gather`` suffices."""
    import torch

    d = _dist()
    rank, ws = world()
    if d is None or ws == 1:
        return [np.asarray(local[i]) for i in range(n_units)]
    mine = shard_units(n_units, rank, ws)
    assert sorted(local) == mine, "local results do not match this rank's shard"
    per_rank = -(-n_units // ws)
    sample = np.asarray(local[mine[0]]) if mine else None
    # ranks with an empty shard learn shape/dtype from rank 0 (which always owns unit 0)
    meta = [sample.shape, str(sample.dtype)] if rank == 0 else None
    box = [meta]
    d.broadcast_object_list(box, src=0)
    shape, dtype = tuple(box[0][0]), np.dtype(box[0][1])
    backend = d.get_backend()
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    is_complex = np.issubdtype(dtype, np.complexfloating)
    real_dtype = np.dtype(dtype.char.lower()) if is_complex else dtype
    width = int(np.prod(shape)) * (2 if is_complex else 1)
    buf = np.zeros((per_rank, width), dtype=real_dtype)
    for j, i in enumerate(mine):
        buf[j] = np.ascontiguousarray(local[i]).view(real_dtype).reshape(-1)
    send = torch.from_numpy(buf).to(device)
    recv = [torch.empty_like(send) for _ in range(ws)]
    d.all_gather(recv, send)
    out = [None] * n_units
    for r in range(ws):
        block = recv[r].cpu().numpy()
        for j, i in enumerate(shard_units(n_units, r, ws)):
            row = block[j]
            out[i] = (row.view(dtype) if is_complex else row).reshape(shape).copy()
    return out


_worker_streams: dict = {}
_pool_lock = None  # created on first use: one run_concurrent at a time per process (the worker streams are shared)


MAX_WORKERS = 8


def balanced_workers(n_units: int, max_workers: int = MAX_WORKERS) -> int:
    """Fewest workers that finish ``n_units`` equal units in the minimal number of rounds (11 units, at most 8 in flight:
    two rounds either way, so 6 workers — less contention per unit than 8)."""
    if n_units <= 0:
        return 1
    rounds = -(-n_units // max(1, max_workers))
    return -(-n_units // rounds)


def run_concurrent(fn, units, workers: int | None = None, cuda: bool = True):
    """Run ``fn(unit)`` for every unit of this rank with up to ``workers`` units in flight on ONE GPU: one host thread
    and one CUDA stream per worker, so that the kernels of independent small waveforms (N <= 2^18: a step-loop launch
    fills a fraction of the 148 SMs and is a dependent chain through a few warps per SM) overlap on the device.
    Measured on one B200 (profiles/r2_unit_concurrency.jsonl): 64-seed Monte-Carlo sweep at N = 2^18 3.5x faster with 8
    units in flight, 11-channel DBP at N = 2^17 4.2x faster with 6.  ``workers=None``: ``balanced_workers(len(units))``.

    Every worker makes the caller's device current, runs its units inside ``torch.cuda.stream(own_stream)`` — this
    package's entry points launch on the current stream and keep one plan / workspace per stream — and the streams are
    ordered against the caller's stream by events on both sides (no host synchronisation): work enqueued by the caller
    before the call is visible to the workers, and the caller's stream waits for all of them before it continues.
    Returns ``{unit: result}``.  The first exception of a worker is re-raised after all workers have stopped.
    One call at a time per process (a second caller waits); ``fn`` must not call ``run_concurrent`` itself.
    ``cuda=False`` runs the same scheduler on plain host threads (tests of the host logic)."""
    import queue
    import threading

    units = list(units)
    out, errors = {}, []
    if not units:
        return out
    workers = balanced_workers(len(units)) if workers is None else max(1, min(int(workers), len(units)))
    if cuda:
        import torch

        dev = torch.cuda.current_device()
        caller = torch.cuda.current_stream(dev)
        ready = torch.cuda.Event()
        ready.record(caller)
        # the same worker streams on every call: plans and cuFFT handles are cached per stream
        pool = _worker_streams.setdefault(dev, [])
        while len(pool) < workers:
            pool.append(torch.cuda.Stream(device=dev))
        streams = pool[:workers]
    q = queue.SimpleQueue()
    for u in units:
        q.put(u)

    def mark(x, stream):  # results allocated on a worker stream are handed to the caller's stream
        import torch

        if isinstance(x, torch.Tensor):
            if x.is_cuda:
                x.record_stream(stream)
        elif isinstance(x, (tuple, list)):
            for y in x:
                mark(y, stream)
        elif isinstance(x, dict):
            for y in x.values():
                mark(y, stream)

    def work(w):
        try:
            if cuda:
                torch.cuda.set_device(dev)  # a new thread starts on device 0
                streams[w].wait_event(ready)
            while not errors:
                try:
                    u = q.get_nowait()
                except queue.Empty:
                    return
                if cuda:
                    with torch.cuda.stream(streams[w]):
                        r = fn(u)
                    mark(r, caller)
                else:
                    r = fn(u)
                out[u] = r
        except BaseException as e:  # noqa: BLE001 - re-raised in the caller
            errors.append(e)

    global _pool_lock
    if _pool_lock is None:
        _pool_lock = threading.Lock()
    threads = [threading.Thread(target=work, args=(w,), name=f"ocb-unit-worker-{w}") for w in range(workers)]
    with _pool_lock:  # a second caller (another host thread) waits: the cached worker streams and their plans are not shared
        for t in threads:
            t.start()
        for t in threads:
            t.join()
    if cuda:
        for s in streams:
            caller.wait_stream(s)
    if errors:
        raise errors[0]
    return out


def run_sharded(fn, units, *args, in_flight: int = 1, **kwargs):
    """Apply ``fn(unit, *args, **kwargs) -> ndarray`` to this rank's shard and gather everything.  ``in_flight`` > 1 keeps
    that many of the rank's units in flight on its GPU (``run_concurrent``); 1 runs them one after the other."""
    mine = shard_units(len(units))
    if in_flight > 1 and len(mine) > 1:
        import torch

        # the scheduler is the same on a CPU-only rank (gloo tests of the host logic): plain host threads, no streams
        local = run_concurrent(lambda i: fn(units[i], *args, **kwargs), mine, in_flight, cuda=torch.cuda.is_available())
    else:
        local = {i: fn(units[i], *args, **kwargs) for i in mine}
    return gather_results(local, len(units))


def gather_device(local: dict, n_units: int):
    """All-gather per-unit CUDA tensors WITHOUT a host round trip: ``local`` maps unit index -> tensor (same shape and
    dtype for every unit; complex allowed).  Ragged ownership is padded to the largest shard; one ``all_gather`` over
    NCCL / NVLink.  Returns the list of ``n_units`` tensors (views into the gathered buffers) on every rank."""
    import torch

    d = _dist()
    rank, ws = world()
    mine = shard_units(n_units, rank, ws)
    assert sorted(local) == mine, "local results do not match this rank's shard"
    if d is None or ws == 1:
        return [local[i] for i in range(n_units)]
    per_rank = -(-n_units // ws)
    sample = local[mine[0]] if mine else None
    meta = [tuple(sample.shape), str(sample.dtype)] if rank == 0 else None
    box = [meta]
    d.broadcast_object_list(box, src=0)
    shape, dtype = box[0][0], getattr(torch, box[0][1].split(".")[-1])
    if sample is not None:
        dev = sample.device
    else:
        dev = torch.device("cuda", torch.cuda.current_device()) if d.get_backend() == "nccl" else torch.device("cpu")
    send = torch.zeros((per_rank,) + tuple(shape), dtype=dtype, device=dev)
    for j, i in enumerate(mine):
        send[j].copy_(local[i])
    is_complex = send.is_complex()
    flat = torch.view_as_real(send) if is_complex else send
    recv = [torch.empty_like(flat) for _ in range(ws)]
    d.all_gather(recv, flat.contiguous())
    out = [None] * n_units
    for r in range(ws):
        block = torch.view_as_complex(recv[r]) if is_complex else recv[r]
        for j, i in enumerate(shard_units(n_units, r, ws)):
            out[i] = block[j]
    return out
