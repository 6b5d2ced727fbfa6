"""BASELINE configs 4 and 5 expressed with the sharding layer (one process per GPU, no data-path collective, ONE
all_gather at the end).  Launch: `python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr
127.0.0.1 tools/example_sharded.py [--case dbp|mc] [--samples 65536]` (or plain `python` for one GPU).

  dbp : digital back-propagation of each WDM-channel-like waveform of a list, one channel per shard unit (cfg4:
        11 channels over 8 GPUs -> shards of 2,2,2,1,1,1,1,1).
  mc  : Monte-Carlo sweep over EDFA noise seeds of one waveform (cfg5: 64 seeds, 8 per GPU); every unit returns
        three scalars (signal power, noise power, OSNR-like ratio), so the gather moves 24 bytes per seed.

The `--dry` flag replaces the GPU calls by a numpy stand-in so that the sharding / gather logic can be exercised
on a CPU box with the gloo backend; it is not a compute path of the package.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from opticommpy_b200.sharding import run_sharded, shard_units, world
from opticommpy_b200.utils import parameters


def fiber_param(seed=None, **kw):
    p = parameters()
    p.Fs, p.Ltotal, p.Lspan, p.hz = 64e9, 160, 80, 2.0
    p.alpha, p.D, p.gamma, p.Fc = 0.2, 16, 1.3, 193.1e12
    p.amp, p.NF, p.maxIter, p.tol, p.nlprMethod = "edfa", 4.5, 10, 1e-5, False
    p.saveSpanN, p.prgsBar, p.seed = [], False, seed
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", choices=["dbp", "mc"], default="mc")
    ap.add_argument("--samples", dest="n", type=int, default=1 << 16)
    ap.add_argument("--units", type=int, default=0)
    ap.add_argument("--dry", action="store_true")
    a = ap.parse_args()
    if "RANK" in os.environ:
        dist.init_process_group("gloo" if a.dry or not torch.cuda.is_available() else "nccl")
        if torch.cuda.is_available() and not a.dry:
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    rank, ws = world()
    rng = np.random.default_rng(0)  # identical unit list on every rank
    if a.dry:
        propagate = lambda x, p: x * np.exp(-0.5j)                       # noqa: E731  (stand-in, CPU logic test only)
        backprop = lambda x, p: x * np.exp(+0.5j)                        # noqa: E731
    else:
        from opticommpy_b200.channels import manakovSSF as propagate
        from opticommpy_b200.equalization import manakovDBP as backprop

    if a.case == "dbp":
        n_units = a.units or 11
        chans = [(rng.normal(size=(a.n, 2)) + 1j * rng.normal(size=(a.n, 2))) * np.sqrt(5e-4) for _ in range(n_units)]

        def unit(x):
            y = propagate(x, fiber_param(amp="ideal"))
            return backprop(y, fiber_param(amp="ideal")).astype(np.complex64)

        out = run_sharded(unit, chans)
        err = [float(np.linalg.norm(o - x) / np.linalg.norm(x)) for o, x in zip(out, chans)]
        if rank == 0:
            print({"case": "dbp", "world": ws, "shard_sizes": [len(shard_units(n_units, r, ws)) for r in range(ws)],
                   "max_round_trip_rel_l2": max(err)})
    else:
        n_units = a.units or 64
        x = (rng.normal(size=(a.n, 2)) + 1j * rng.normal(size=(a.n, 2))) * np.sqrt(5e-4)
        clean = propagate(x, fiber_param(amp="ideal"))

        def unit(seed):
            y = propagate(x, fiber_param(seed=int(seed), noiseRNG="philox"))
            ps, pn = np.mean(np.abs(clean) ** 2), np.mean(np.abs(y - clean) ** 2)
            return np.array([ps, pn, 10 * np.log10(ps / max(pn, 1e-300))])

        out = np.stack(run_sharded(unit, list(range(1000, 1000 + n_units))))
        if rank == 0:
            print({"case": "mc", "world": ws, "seeds": n_units, "gathered_shape": out.shape,
                   "snr_db_mean": float(out[:, 2].mean()), "snr_db_std": float(out[:, 2].std())})
    if dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
