#!/usr/bin/env python
"""One profiled Primal-CR++ outer iteration (for ncu): builds the workload, runs `--warmup` iterations, then brackets
ONE outer iteration with cudaProfilerStart/Stop (use `ncu --profile-from-start off`).  Never a bench number."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="netflix")
    ap.add_argument("--scale", type=float, default=0.1)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--lam", type=float, default=5000.0)
    ap.add_argument("--warmup", type=int, default=1)
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
    for _ in range(args.warmup):
        e.outer_iteration()
    torch.cuda.synchronize()
    rt = torch.cuda.cudart()
    rt.cudaProfilerStart()
    t = time.time()
    obj = e.outer_iteration()
    torch.cuda.synchronize()
    dt = time.time() - t
    rt.cudaProfilerStop()
    print("profiled outer iteration: nnz=%d obj=%.6g wall=%.3fs counters=%s" % (ds.train.nnz, obj, dt, e.counters()))
    e.close()


if __name__ == "__main__":
    main()
