#!/usr/bin/env python
"""Runs the other BASELINE.json configs at a chosen scale on one GPU and checks size-independent properties:
objective decreases, the objective reported by update_U equals the objective recomputed from scratch, Primal-CR and
Primal-CR++ agree on it.  Prints one JSON line per shape (timings are informative, not bench numbers)."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run(shape, scale, k, solver, iters, lam=5000.0, check_pcr=False):
    import torch
    from primalcr_b200 import api
    from primalcr_b200.data import synth_dataset
    t = time.time()
    ds = synth_dataset(shape, scale=scale, device="cuda", test_per_user=0)
    torch.cuda.empty_cache()
    gen = time.time() - t
    U = api.reference_init(ds.d1, k)
    V = U[:ds.d2].copy() if ds.d2 <= ds.d1 else api.reference_init(ds.d2, k)
    e = api.Engine(api.Parameter(solver_type=solver, k=k, lambda_=lam, maxiter=iters, do_predict=0))
    e.set_levels(np.arange(1, 6))
    t = time.time(); e.set_train(ds.train); e.set_factors(U, V); setup = time.time() - t
    objs = [e.initial_objective()]
    times, counters = [], []
    for it in range(iters):
        if it == iters - 1:
            e.profile_enable(True); e.profile_reset()
        t = time.time(); objs.append(e.outer_iteration()); times.append(time.time() - t); counters.append(e.counters())
    prof = e.profile(); e.profile_enable(False)
    top = sorted(((v["ms"], n, v["launches"]) for n, v in prof.items()), reverse=True)[:12]
    recomputed = e.initial_objective()
    out = dict(shape=shape, scale=scale, solver=solver, k=k, d1=ds.d1, d2=ds.d2, nnz=ds.train.nnz,
               max_len=int(ds.train.lens().max()), gen_s=gen, setup_s=setup, sec_per_iter=times, objective=objs,
               recomputed_rel_err=abs(recomputed - objs[-1]) / abs(objs[-1]), monotone=bool(np.all(np.diff(objs) < 0)),
               device_gb=e.device_bytes() / 1e9, counters=counters[-1],
               kernels_last_iter=[dict(name=n, ms=round(ms, 2), launches=l) for ms, n, l in top])
    if check_pcr:
        Ug, Vg = e.get_factors()
        e1 = api.Engine(api.Parameter(solver_type=3 - solver, k=k, lambda_=lam, maxiter=1, do_predict=0))
        e1.set_levels(np.arange(1, 6)); e1.set_train(ds.train); e1.set_factors(Ug, Vg)
        o = e1.initial_objective()
        out["other_solver_rel_err"] = abs(o - objs[-1]) / abs(objs[-1])
        e1.close()
    err, ndcg = e.eval(0)
    out["train_pairwise_error"] = err; out["train_ndcg"] = ndcg
    e.close()
    print(json.dumps(out), flush=True)
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--which", default="ml1m_pcr,ml1m_pcrpp,yahoo,powerlaw")
    ap.add_argument("--yahoo-scale", type=float, default=0.2)
    ap.add_argument("--powerlaw-scale", type=float, default=0.05)
    a = ap.parse_args()
    w = a.which.split(",")
    if "ml1m_pcrpp" in w:
        run("ml1m", 1.0, 10, 2, 3, check_pcr=True)             # BASELINE config #1 shape
    if "ml1m_pcr" in w:
        run("ml1m", 1.0, 100, 1, 2, check_pcr=True)            # BASELINE config #2: Primal-CR k=100
    if "yahoo" in w:
        run("yahoo", a.yahoo_scale, 100, 2, 2)                 # config #4 shape (V not L2-resident)
    if "powerlaw" in w:
        run("powerlaw", a.powerlaw_scale, 200, 2, 2)           # config #5 shape (heavy users, k=200)
