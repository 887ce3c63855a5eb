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


def _worker_sharded_cg(rank, world, port, out):
    """The sharded form of solve_delta_new (engine.cu update_V, large item sets): every rank holds its users' partial Hp
    over ALL items; reduce-scatter by item rows, CG recurrences on the rank's row slice with the dot products summed over
    ranks, all-gather of the search direction p (and of delta at the end).  Same control flow as pcrpp.cpp:335-358."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ds = synth_dataset("tiny")
    k, lam = 5, 30.0
    U, V = np_init(ds.d1, ds.d2, k, seed=3, scale=0.5)
    b = shard_bounds(ds.train.row_ptr, world)
    u0, u1 = int(b[rank]), int(b[rank + 1])
    X = to_csr(ds.train.slice_users(u0, u1)); Ul = np.ascontiguousarray(U[u0:u1])
    O = ob.oracle()
    m = O.comp_m(X, Ul, V)
    d2p = (ds.d2 + world - 1) // world * world                       # rows padded to a multiple of the world size
    rows = d2p // world; r0 = rank * rows

    def pad(M):
        P = np.zeros((d2p, k)); P[:ds.d2] = M; return P

    def reduce_scatter_rows(partial_full, x_full):                   # -> my row slice of sum_r partial_r + lambda x
        t = torch.from_numpy(pad(partial_full))
        dist.all_reduce(t)                                           # gloo has no reduce_scatter: all_reduce + slice, same sums
        return t[r0:r0 + rows].numpy().copy() + lam * pad(x_full)[r0:r0 + rows]

    def all_gather_rows(slice_):
        parts = [torch.zeros(rows, k, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(parts, torch.from_numpy(np.ascontiguousarray(slice_)))
        return torch.cat(parts).numpy()[:ds.d2]

    def sdot(a, c):
        t = torch.tensor([float((a * c).sum())], dtype=torch.float64); dist.all_reduce(t); return float(t)

    g_s = reduce_scatter_rows(O.obtain_g_new(X, Ul, V, m, lam) - lam * V, V)
    delta_s = np.zeros_like(g_s); rr_s = -g_s; p_s = g_s.copy()
    p = all_gather_rows(p_s)
    err = 0.01 * np.sqrt(sdot(rr_s, rr_s))
    its = 0
    for _ in range(10):
        Hp_s = reduce_scatter_rows(O.compute_Ha_new(X, p, m, Ul, lam) - lam * p, p)
        its += 1
        pHp = sdot(p_s, Hp_s); alpha = -sdot(rr_s, p_s) / pHp
        delta_s = delta_s + alpha * p_s; rr_s = rr_s + alpha * Hp_s
        if np.sqrt(sdot(rr_s, rr_s)) < err:
            break
        beta = sdot(rr_s, Hp_s) / pHp
        p_s = -rr_s + beta * p_s
        p = all_gather_rows(p_s)
    delta = all_gather_rows(delta_s)
    if rank == 0:
        np.savez(out, delta=delta, its=its)
    dist.destroy_process_group()


def test_two_rank_sharded_cg_matches_single_process(tmp_path):
    out = str(tmp_path / "cg.npz")
    mp.spawn(_worker_sharded_cg, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    ds = synth_dataset("tiny")
    k, lam = 5, 30.0
    U, V = np_init(ds.d1, ds.d2, k, seed=3, scale=0.5)
    O = ob.oracle(); X = to_csr(ds.train)
    m = O.comp_m(X, U, V)
    g = O.obtain_g_new(X, U, V, m, lam)
    delta = np.zeros_like(g); rr = -g; p = g.copy(); err = 0.01 * np.sqrt((rr * rr).sum()); its = 0
    for _ in range(10):                                               # solve_delta_new pcrpp.cpp:335-358, unsharded
        Hp = O.compute_Ha_new(X, p, m, U, lam); its += 1
        pHp = (p * Hp).sum(); alpha = -(rr * p).sum() / pHp
        delta = delta + alpha * p; rr = rr + alpha * Hp
        if np.sqrt((rr * rr).sum()) < err:
            break
        p = -rr + ((rr * Hp).sum() / pHp) * p
    assert int(got["its"]) == its
    assert np.abs(got["delta"] - delta).max() <= 1e-10 * np.abs(delta).max()
