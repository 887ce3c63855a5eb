#!/usr/bin/env python
"""Times the training-set / test-set evaluation (pairwise error + NDCG@10) at a chosen scale (informative only)."""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="netflix"); ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--k", type=int, default=100)
    a = ap.parse_args()
    import torch
    from primalcr_b200 import api
    from primalcr_b200.data import synth_dataset
    ds = synth_dataset(a.workload, scale=a.scale, device="cuda", test_per_user=10)
    torch.cuda.empty_cache()
    U = api.reference_init(ds.d1, a.k); V = U[:ds.d2].copy() if ds.d2 <= ds.d1 else api.reference_init(ds.d2, a.k)
    e = api.Engine(api.Parameter(solver_type=2, k=a.k, lambda_=5000.0, maxiter=1, do_predict=1))
    e.set_levels(np.arange(1, 6)); e.set_train(ds.train); e.set_test(ds.test); e.set_factors(U, V)
    e.initial_objective(); e.outer_iteration()
    torch.cuda.synchronize(); t = time.time(); e.outer_iteration(); torch.cuda.synchronize()
    out = {"outer_iteration_sec": time.time() - t}
    # method -1: what primalcr_eval picks (sorted-state count for the Primal-CR++ training set), 0: all pairs, 1: sorted state
    for which, name, method in ((0, "train_auto", -1), (0, "train_all_pairs", 0), (0, "train_sorted", 1), (1, "test", -1)):
        e.profile_enable(True); e.profile_reset()
        torch.cuda.synchronize(); t = time.time()
        r = e.eval(which) if method < 0 else e.eval_error_counts(which, method)[:2]
        torch.cuda.synchronize(); dt = time.time() - t
        prof = e.profile(); e.profile_enable(False)
        out[name] = dict(sec=dt, result=r, kernels={n: round(v["ms"], 3) for n, v in prof.items() if v["ms"] > 0.01})
    print(json.dumps(dict(workload=a.workload, scale=a.scale, nnz=ds.train.nnz, max_len=int(ds.train.lens().max()), **out)))
    e.close()

if __name__ == "__main__":
    main()
