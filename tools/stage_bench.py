#!/usr/bin/env python
"""Stage-level timing helper (never a bench number): the V-side and U-side Hessian-vector stages (dots -> sweep -> row sum)
called `--reps` times through the C ABI at the named shape, per-kernel CUDA-event table printed as one JSON line.
Used for A/B switches that would derail a real trajectory (e.g. PRIMALCR_EXP_* timing experiments)."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="netflix")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--lam", type=float, default=5000.0)
    ap.add_argument("--reps", type=int, default=4)
    ap.add_argument("--warm-iters", type=int, default=1)
    ap.add_argument("--side", default="V", choices=["V", "U", "VU"])
    ap.add_argument("--tag", default="")
    args = ap.parse_args()
    import torch
    from primalcr_b200 import api
    from primalcr_b200.data import synth_dataset
    ds = synth_dataset(args.workload, scale=args.scale, device="cuda", test_per_user=0)
    torch.cuda.empty_cache()
    U = api.reference_init(ds.d1, args.k)
    V = U[:ds.d2].copy() if ds.d2 <= ds.d1 else api.reference_init(ds.d2, args.k)
    e = api.Engine(api.Parameter(solver_type=2, k=args.k, lambda_=args.lam, maxiter=1, do_predict=0))
    e.set_levels(np.arange(1, 6))
    e.set_train(ds.train); e.set_factors(U, V)
    e.initial_objective()
    for _ in range(args.warm_iters):
        e.outer_iteration()
    a = np.random.default_rng(1).standard_normal((ds.d2, args.k))
    S = np.random.default_rng(2).standard_normal((ds.d1, args.k)) if "U" in args.side else None
    if "V" in args.side:
        e.hv_V(a)
    if "U" in args.side:
        e.hv_U(S)
    e.profile_enable(True); e.profile_reset()
    for _ in range(args.reps):
        if "V" in args.side:
            e.hv_V(a)
        if "U" in args.side:
            e.hv_U(S)
    prof = e.profile(); e.profile_enable(False)
    kern = sorted(((v["ms"] / max(v["launches"], 1), n, v["launches"]) for n, v in prof.items()), reverse=True)
    out = dict(tag=args.tag, env={k: v for k, v in os.environ.items() if k.startswith("PRIMALCR_")}, workload=args.workload,
               scale=args.scale, k=args.k, nnz=ds.train.nnz, ms_per_launch={n: [round(ms, 4), l] for ms, n, l in kern[:16]})
    e.close()
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
