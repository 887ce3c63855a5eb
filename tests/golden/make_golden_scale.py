#!/usr/bin/env python
"""Regenerates tests/golden/scale_*.npz: parity fixtures at the sizes the headline is quoted on.
Run in the build container only (needs oracle/_ref):  python tests/golden/make_golden_scale.py [case ...]

Cases (data sets come from the committed deterministic generator, CPU stream; the fixture stores a checksum of the CSR)
  netflix005   Netflix-shape x 0.05 (24,009 users x 17,770 items, 5.0 M ratings, 13 users above 4096 ratings),
               Primal-CR++ k=100 lambda=5000, 2 outer iterations, evaluation on
  powerlaw001  power-law x 0.01 (20,000 users x 500,000 items, 5.0 M ratings, one 99,990-rating user, 120 heavy users),
               Primal-CR++ k=200 lambda=5000, 2 outer iterations, evaluation on
  ml1m_pcr     ml1m-shape (6040 x 3952, 939,809 ratings), Primal-CR (-s 1) k=100 lambda=5000, 2 outer iterations
Each fixture holds, from the single-threaded C restatement (bit-identical to `omp-pmf-train -n 1`, tests/test_oracle.py):
objective per iteration, the 8 control-flow counters per iteration, error / NDCG@10 per iteration, the integer pair-error
count of every user (training and test set) and the users whose row has collapsed to rounding noise after every
iteration, sampled rows and column sums of the final U and V; and from the UNMODIFIED reference (race-free objects, all host threads) the same
objectives and evaluation numbers, so the fixture itself shows oracle == reference at this size.  powerlaw001 also
carries `ref1_*`: the unmodified reference with ONE thread (21 minutes), which the restatement matches bit for bit
(objective, evaluation, sampled rows of U and V) -- while the all-threads run of the same reference differs from it by
7e-7 .. 3e-6 of the objective at iteration 2 from run to run (the order of its `omp atomic` adds): that spread is the
tolerance the GPU test uses for this one case.
"""
import hashlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bindings as ob                      # noqa: E402
from primalcr_b200 import api                          # noqa: E402  (reference_init: host code, no GPU)
from primalcr_b200.data import synth_dataset           # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    "netflix005": dict(shape="netflix", scale=0.05, solver=2, k=100, lam=5000.0, iters=2, predict=1),
    "powerlaw001": dict(shape="powerlaw", scale=0.01, solver=2, k=200, lam=5000.0, iters=2, predict=1),
    "ml1m_pcr": dict(shape="ml1m", scale=1.0, solver=1, k=100, lam=5000.0, iters=2, predict=1),
}


def csr(R):
    return ob.Csr(R.d1, R.d2, R.row_ptr, R.item.astype(np.int64), R.rating)


def dataset_digest(ds):
    h = hashlib.sha1()
    for a in (ds.train.row_ptr, ds.train.item, ds.train.rating, ds.test.row_ptr, ds.test.item, ds.test.rating):
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def factor_summary(M, rows):
    return dict(rows=M[rows].copy(), colsum=M.sum(0), sq=float((M * M).sum()), absmax=float(np.abs(M).max()))


def make(name):
    c = CASES[name]
    t = time.time()
    ds = synth_dataset(c["shape"], scale=c["scale"], device="cpu")
    k, lam, iters, solver = c["k"], c["lam"], c["iters"], c["solver"]
    X, XT = csr(ds.train), csr(ds.test)
    U0 = api.reference_init(ds.d1, k); V0 = api.reference_init(ds.d2, k)
    print(name, "data %.0fs: d1=%d d2=%d nnz=%d test=%d maxlen=%d" % (time.time() - t, ds.d1, ds.d2, ds.train.nnz, ds.test.nnz,
                                                                    int(ds.train.lens().max())), flush=True)
    t = time.time()
    R = ob.reference_rf()
    assert R is not None, "build oracle/_ref first (make -C oracle ref)"
    ref = R.train(solver, X, XT, U0, V0, lam, iters, do_predict=c["predict"], threads=os.cpu_count() or 1)
    print(name, "reference (race-free, %d threads) %.0fs obj %s" % (os.cpu_count(), time.time() - t, ref["obj"]), flush=True)
    # single-threaded C restatement, ONE outer iteration per call so that the state after every iteration is visible:
    # objective, counters, evaluation incl. the integer pair-error count of every user, and the set of users whose row
    # collapsed to rounding noise (zero loss gradient => the Newton step returns u_i - u_i; the ORDER of such a user's
    # ~1e-17 scores, hence its evaluation, is noise in the reference too -- the GPU test masks exactly these users)
    t = time.time()
    O = ob.oracle()
    U, V = U0, V0
    objs, counters, evals, err_train, err_test, collapsed = [], [], [], [], [], []

    def evaluate(U, V):
        row = []
        for Xs, store in ((X, err_train), (XT, err_test)):
            out, _, per_user = O.eval(Xs, U, V, 10, want_counts=True)
            row += [out[0], out[1]]; store.append(per_user.copy())
        evals.append(row); collapsed.append(np.abs(U).max(1) < 1e-9)

    for it in range(iters + 1):
        res = O.train(solver, X, None, U, V, lam, 1 if it else 0, do_predict=0)
        if it == 0:
            objs.append(res["obj"][0])
        else:
            objs.append(res["obj"][1]); counters.append(res["counters"][0]); U, V = res["U"], res["V"]
        evaluate(U, V)
        print(name, "oracle iteration %d: obj %.12g, %d collapsed users (%.0fs)" % (it, objs[-1], collapsed[-1].sum(), time.time() - t), flush=True)
    orc = dict(obj=np.array(objs), counters=np.array(counters), evals=np.array(evals), U=U, V=V)
    rel = np.abs(orc["obj"] - ref["obj"]) / np.abs(ref["obj"])
    print(name, "oracle vs reference: objective rel err", rel, "evals abs err", np.abs(orc["evals"] - ref["evals"]).max(),
          "U", np.abs(orc["U"] - ref["U"]).max() / np.abs(ref["U"]).max(), "V", np.abs(orc["V"] - ref["V"]).max() / np.abs(ref["V"]).max(),
          flush=True)
    rng = np.random.default_rng(99)
    urows = np.sort(rng.choice(ds.d1, size=min(256, ds.d1), replace=False))
    vrows = np.sort(rng.choice(ds.d2, size=min(256, ds.d2), replace=False))
    heavy = np.argsort(ds.train.lens())[-8:]
    urows = np.unique(np.concatenate([urows, heavy]))
    out = dict(shape=c["shape"], scale=c["scale"], solver=solver, k=k, lam=lam, iters=iters, predict=c["predict"],
               d1=ds.d1, d2=ds.d2, nnz=ds.train.nnz, nnz_test=ds.test.nnz, digest=dataset_digest(ds),
               obj=orc["obj"], evals=orc["evals"], counters=orc["counters"],
               ref_obj=ref["obj"], ref_evals=ref["evals"], ref_threads=os.cpu_count() or 1,
               err_train=np.array(err_train), err_test=np.array(err_test), collapsed=np.packbits(np.array(collapsed), axis=1),
               urows=urows, vrows=vrows)
    for tag, res in (("", orc), ("ref_", ref)):
        for nm, M, rows in (("U", res["U"], urows), ("V", res["V"], vrows)):
            s = factor_summary(M, rows)
            out[tag + nm + "_rows"] = s["rows"]; out[tag + nm + "_colsum"] = s["colsum"]
            out[tag + nm + "_sq"] = s["sq"]; out[tag + nm + "_absmax"] = s["absmax"]
    np.savez_compressed(os.path.join(HERE, "scale_%s.npz" % name), **out)
    print(name, "written", flush=True)


if __name__ == "__main__":
    for n in (sys.argv[1:] or list(CASES)):
        make(n)
