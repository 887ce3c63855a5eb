#!/usr/bin/env python
"""Regenerates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref, built from /root/reference by
oracle/Makefile).  Run in the build container only:  python tests/golden/make_golden.py

Fixtures
  golden_tiny.npz     synthetic 'tiny' shape (300 x 120), integer ratings 1-5, train + test
  golden_ml1m600.npz  first 600 users of the reference's bundled ml1m/test.ratings used as a TRAINING set
                      (10 ratings per user, integers 1-5), no test set
  golden_toy400.npz   first 400 users of the reference's bundled toy-example/test.ratings as a training set:
                      real-valued ratings (lround gives 9 levels; Primal-CR compares exact doubles); its "test set" is
                      the same ratings minus the users that degenerate to u_i ~ 1e-17 after some iteration (see main())
Each holds the CSR arrays, the reference init, and for solver 1 and 2 the outputs of the reference driver loop
(objective per iteration at full precision, pairwise error / NDCG@10 per iteration, final U and V), the stage
outputs (scores, g, Ha, objective) at the initial point, plus the 6-digit stdout of `omp-pmf-train -n 1`.
"""
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bindings as ob                      # noqa: E402
from primalcr_b200.data import (Dataset, Ratings, csr_in_file_order, read_ratings_file, synth_dataset,  # noqa: E402
                                write_reference_dir)

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def csr(R):
    return ob.Csr(R.d1, R.d2, R.row_ptr, R.item.astype(np.int64), R.rating)


def head_users(path, n_users, d2):
    u, i, r = read_ratings_file(path)
    keep = u < n_users
    items, compact = np.unique(i[keep], return_inverse=True)     # relabel to the items present (keeps fixtures small)
    return Ratings.from_coo(n_users, len(items), u[keep], compact, r[keep])


def cli_stdout(ds, solver, k, lam, iters):
    exe = ob.ref_cli("omp-pmf-train")
    with tempfile.TemporaryDirectory() as tmp:
        d = os.path.join(tmp, "data"); write_reference_dir(d, ds)
        out = subprocess.run([exe, "-s", str(solver), "-k", str(k), "-l", str(lam), "-t", str(iters), "-p", "1", "-n", "1",
                              d, os.path.join(tmp, "m")], cwd=tmp, capture_output=True, text=True, check=True).stdout
    return re.sub(r"time \S+", "time T", re.sub(r"Wall-time: \S+", "Wall-time: T", out))


def make(name, ds, k, lam, iters):
    R = ob.reference()
    assert R is not None, "build oracle/_ref first (make -C oracle ref)"
    X, XT = csr(ds.train), csr(ds.test)
    U0 = ob.ref_initial(ds.d1, k); V0 = ob.ref_initial(ds.d2, k)
    out = dict(d1=ds.d1, d2=ds.d2, k=k, lam=lam, iters=iters,
               row_ptr=ds.train.row_ptr, item=ds.train.item, rating=ds.train.rating,
               t_row_ptr=ds.test.row_ptr, t_item=ds.test.item, t_rating=ds.test.rating, U0=U0, V0=V0)
    m = R.comp_m(X, U0, V0)
    a = np.random.default_rng(1234).standard_normal(V0.shape)
    out.update(m0=m, dir_a=a,
               g2=R.obtain_g_new(X, U0, V0, m, lam), Ha2=R.compute_Ha_new(X, a, m, U0, lam),
               obj2=R.objective_new(X, m, U0, V0, lam),
               g1=R.pcr_obtain_g(X, U0, V0, m, lam), Ha1=R.pcr_compute_Ha(X, a, m, U0, lam),
               obj1=R.pcr_objective(X, m, U0, V0, lam),
               eval_train0=R.eval(X, U0, V0, 10))
    if ds.test.nnz:
        out["eval_test0"] = R.eval(XT, U0, V0, 10)
    for solver in (1, 2):
        res = R.train(solver, X, XT, U0, V0, lam, iters, do_predict=1)
        out["s%d_obj" % solver] = res["obj"]; out["s%d_evals" % solver] = res["evals"]
        out["s%d_U" % solver] = res["U"]; out["s%d_V" % solver] = res["V"]
        out["s%d_stdout" % solver] = np.array(cli_stdout(ds, solver, k, lam, iters))
    np.savez_compressed(os.path.join(HERE, "golden_%s.npz" % name), **out)
    print(name, "written:", {k_: (v.shape if hasattr(v, "shape") else v) for k_, v in out.items() if k_.endswith("obj")})


def main():
    make("tiny", synth_dataset("tiny"), k=7, lam=50.0, iters=3)
    ml = head_users(os.path.join(REF, "ml1m", "test.ratings"), 600, 3952)
    make("ml1m600", Dataset(ml, Ratings.empty(600, ml.d2)), k=10, lam=100.0, iters=3)
    toy = head_users(os.path.join(REF, "toy-example", "test.ratings"), 400, 3952)
    # toy400's "test set" = the training ratings of the users that do NOT degenerate: users whose ratings all round to one
    # level (or whose pairs are separated by the margin from the start) have a zero loss gradient, the Newton step sends
    # u_i to ~1e-17 and the ORDER of their scores is rounding noise -- in the reference too.  The all-users training
    # numbers can only be compared loosely; the masked ones (evals[:, 2:4]) are held to the normal tolerances.
    R = ob.reference()
    U0 = ob.ref_initial(toy.d1, 8); V0 = ob.ref_initial(toy.d2, 8)
    alive = np.ones(toy.d1, bool)
    for solver in (1, 2):
        for it in (1, 2, 3):        # a user can collapse after one iteration and recover in the next: mask the union
            probe = R.train(solver, csr(toy), None, U0, V0, 20.0, it, do_predict=0)
            alive &= np.abs(probe["U"]).max(1) > 1e-9
    keep = np.repeat(alive, toy.lens())
    masked = csr_in_file_order(toy.d1, toy.d2, toy.users()[keep], toy.item[keep], toy.rating[keep])
    print("toy400: %d of %d users degenerate at some iteration (masked out of the test-set evaluation)" % ((~alive).sum(), toy.d1))
    make("toy400", Dataset(toy, masked), k=8, lam=20.0, iters=3)


if __name__ == "__main__":
    main()
