"""world_size-2 gloo test (CPU) of the N>1 host logic: nnz-balanced user shards + sum of per-shard V-side
quantities reproduces the single-process result.  The per-shard arithmetic is the CPU oracle's; what is
under test is the decomposition bench.py / the engines rely on: g = lambda*V + sum_r G_r, obj = sum_r loss_r + reg."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import bindings as ob
from primalcr_b200.data import shard_bounds, synth_dataset
from tests.util import np_init, to_csr


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ds = synth_dataset("tiny")
    k, lam = 6, 30.0
    U, V = np_init(ds.d1, ds.d2, k, seed=9, scale=0.5)
    b = shard_bounds(ds.train.row_ptr, world)
    u0, u1 = int(b[rank]), int(b[rank + 1])
    sh = ds.train.slice_users(u0, u1)
    O = ob.oracle(); X = to_csr(sh)
    Ul = np.ascontiguousarray(U[u0:u1])
    m = O.comp_m(X, Ul, V)
    G = O.obtain_g_new(X, Ul, V, m, lam) - lam * V                 # this shard's scatter part only
    a = np.random.default_rng(4).standard_normal(V.shape)
    H = O.compute_Ha_new(X, a, m, Ul, lam) - lam * a
    loss = O.objective_new(X, m, Ul, V, lam) - lam * ((Ul ** 2).sum() + (V ** 2).sum()) / 2.0
    t = torch.from_numpy(np.concatenate([G.ravel(), H.ravel(), [loss, (Ul ** 2).sum()]]))
    dist.all_reduce(t)                                               # the V-side exchange step (NCCL on the GPUs)
    t = t.numpy()
    n = V.size
    if rank == 0:
        g = t[:n].reshape(V.shape) + lam * V
        Ha = t[n:2 * n].reshape(V.shape) + lam * a
        obj = t[2 * n] + lam * (t[2 * n + 1] + (V ** 2).sum()) / 2.0
        np.savez(out, g=g, Ha=Ha, obj=obj)
    dist.destroy_process_group()


def test_two_rank_decomposition_matches_single_process(tmp_path):
    out = str(tmp_path / "r.npz")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    ds = synth_dataset("tiny")
    k, lam = 6, 30.0
    U, V = np_init(ds.d1, ds.d2, k, seed=9, scale=0.5)
    O = ob.oracle(); X = to_csr(ds.train)
    m = O.comp_m(X, U, V)
    a = np.random.default_rng(4).standard_normal(V.shape)
    assert np.abs(got["g"] - O.obtain_g_new(X, U, V, m, lam)).max() <= 1e-12 * np.abs(got["g"]).max()
    assert np.abs(got["Ha"] - O.compute_Ha_new(X, a, m, U, lam)).max() <= 1e-12 * np.abs(got["Ha"]).max()
    assert abs(float(got["obj"]) - O.objective_new(X, m, U, V, lam)) <= 1e-12 * float(got["obj"])
