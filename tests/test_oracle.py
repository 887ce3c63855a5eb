"""CPU tests (-m "not gpu"): the plain-C oracle against (a) golden vectors produced by the UNMODIFIED reference
(tests/golden/*.npz, generator tests/golden/make_golden.py) and (b) the reference itself where oracle/_ref is
built (this container).  This is what pins the oracle; the GPU parity tests then compare CUDA against it."""
import os

import numpy as np
import pytest

from oracle import bindings as ob
from tests.util import dataset, np_init, rel, to_csr

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = ["tiny", "ml1m600", "toy400"]


def load_gold(name):
    g = np.load(os.path.join(GOLD, "golden_%s.npz" % name))
    X = ob.Csr(int(g["d1"]), int(g["d2"]), g["row_ptr"], g["item"].astype(np.int64), g["rating"])
    XT = ob.Csr(int(g["d1"]), int(g["d2"]), g["t_row_ptr"], g["t_item"].astype(np.int64), g["t_rating"]) \
        if len(g["t_item"]) else None
    return g, X, XT


@pytest.mark.parametrize("name", NAMES)
def test_oracle_stages_match_reference_golden(name):
    g, X, XT = load_gold(name)
    O = ob.oracle()
    U0, V0, lam = g["U0"], g["V0"], float(g["lam"])
    m = O.comp_m(X, U0, V0)
    assert np.array_equal(m, g["m0"])                                    # same t-ascending sum => bit-exact
    assert rel(O.obtain_g_new(X, U0, V0, m, lam), g["g2"]) < 1e-14
    assert rel(O.compute_Ha_new(X, g["dir_a"], m, U0, lam), g["Ha2"]) < 1e-14
    assert abs(O.objective_new(X, m, U0, V0, lam) - float(g["obj2"])) <= 1e-13 * float(g["obj2"])
    assert rel(O.pcr_obtain_g(X, U0, V0, m, lam), g["g1"]) < 1e-14
    assert rel(O.pcr_compute_Ha(X, g["dir_a"], m, U0, lam), g["Ha1"]) < 1e-14
    assert abs(O.pcr_objective(X, m, U0, V0, lam) - float(g["obj1"])) <= 1e-13 * float(g["obj1"])
    assert np.allclose(O.eval(X, U0, V0, 10), g["eval_train0"], rtol=0, atol=1e-14)
    if XT is not None:
        assert np.allclose(O.eval(XT, U0, V0, 10), g["eval_test0"], rtol=0, atol=1e-14)


@pytest.mark.parametrize("solver", [1, 2])
@pytest.mark.parametrize("name", NAMES)
def test_oracle_trajectory_matches_reference_golden(name, solver):
    g, X, XT = load_gold(name)
    res = ob.oracle().train(solver, X, XT, g["U0"], g["V0"], float(g["lam"]), int(g["iters"]), do_predict=1)
    want = g["s%d_obj" % solver]
    assert np.all(np.abs(res["obj"] - want) <= 1e-12 * np.abs(want))
    assert rel(res["U"], g["s%d_U" % solver]) < 1e-12 and rel(res["V"], g["s%d_V" % solver]) < 1e-12
    assert np.allclose(res["evals"], g["s%d_evals" % solver], rtol=0, atol=1e-12, equal_nan=True)
    # and the reference CLI's own 6-digit log agrees with both
    lines = str(g["s%d_stdout" % solver]).splitlines()
    objs = [float(l.split()[-1]) for l in lines if l.startswith("Iter ")]
    assert np.all(np.abs(np.array(objs) - want) <= 1e-5 * np.abs(want))


def test_reference_init_is_the_cli_stream():
    """initial() util.cpp:80-93: a fresh default-seeded engine per call, so V == U[:d2]."""
    g, _, _ = load_gold("tiny")
    assert np.array_equal(g["V0"], g["U0"][:int(g["d2"])])


def test_objective_pcr_equals_pcrpp_on_integer_ratings():
    ds = dataset("tiny")
    U, V = np_init(ds.d1, ds.d2, 5, seed=3, scale=0.4)
    O = ob.oracle(); X = to_csr(ds.train)
    m = O.comp_m(X, U, V)
    a, b = O.objective_new(X, m, U, V, 10.0), O.pcr_objective(X, m, U, V, 10.0)
    assert abs(a - b) <= 1e-12 * abs(b)
    assert rel(O.obtain_g_new(X, U, V, m, 10.0), O.pcr_obtain_g(X, U, V, m, 10.0)) < 1e-11


def test_level_counts_are_window_counts():
    """The loop-local counters exposed by orc_level_counts obey their definitions (brute force)."""
    rng = np.random.default_rng(0)
    for n in (1, 2, 7, 40):
        mm = np.round(rng.standard_normal(n) * 1.5, 1)        # many exact ties and exact +-1.0 gaps
        vals = rng.integers(1, 5, n).astype(float)
        o = ob.level_counts(mm, vals)
        s, lev = o["s"], o["level"]
        for j in range(n):
            for t in range(o["T"]):
                assert o["cntL"][j, t] == np.sum((s <= s[j] + 1.0) & (lev == t))
                assert o["cntR"][j, t] == np.sum((s >= s[j] - 1.0) & (lev == t))


# ---- against the live reference (only where oracle/_ref was built, i.e. in the build container) -----------------
needs_ref = pytest.mark.skipif(ob.reference() is None, reason="oracle/_ref not built (no /root/reference here)")


@needs_ref
@pytest.mark.parametrize("name,k", [("tiny", 6), ("ragged", 9)])
def test_oracle_vs_live_reference_stages(name, k):
    ds = dataset(name)
    X = to_csr(ds.train)
    O, R = ob.oracle(), ob.reference()
    U, V = np_init(ds.d1, ds.d2, k, seed=1, scale=0.5)
    lam = 25.0
    m = O.comp_m(X, U, V)
    assert np.array_equal(m, R.comp_m(X, U, V))
    a = np.random.default_rng(2).standard_normal(V.shape)
    assert rel(O.obtain_g_new(X, U, V, m, lam), R.obtain_g_new(X, U, V, m, lam)) < 1e-14
    assert rel(O.compute_Ha_new(X, a, m, U, lam), R.compute_Ha_new(X, a, m, U, lam)) < 1e-14
    assert abs(O.objective_new(X, m, U, V, lam) / R.objective_new(X, m, U, V, lam) - 1) < 1e-13
    assert rel(O.pcr_obtain_g(X, U, V, m, lam), R.pcr_obtain_g(X, U, V, m, lam)) < 1e-14
    assert rel(O.pcr_compute_Ha(X, a, m, U, lam), R.pcr_compute_Ha(X, a, m, U, lam)) < 1e-14
    assert abs(O.pcr_objective(X, m, U, V, lam) / R.pcr_objective(X, m, U, V, lam) - 1) < 1e-13
    assert np.allclose(O.eval(X, U, V), R.eval(X, U, V), rtol=0, atol=1e-14)
    assert np.allclose(O.eval(to_csr(ds.test), U, V), R.eval(to_csr(ds.test), U, V), rtol=0, atol=1e-14)
    rp = ds.train.row_ptr
    s = np.random.default_rng(3).standard_normal(k)
    for u in range(0, ds.d1, max(1, ds.d1 // 25)):
        lo, hi = int(rp[u]), int(rp[u + 1])
        go, oo, ho = O.user_stage(X.rows[lo:hi], X.vals[lo:hi], m[lo:hi], V, lam, U[u], s)
        gr, orr, hr = R.user_stage(X.rows[lo:hi], X.vals[lo:hi], m[lo:hi], V, lam, U[u], s)
        assert rel(go, gr) < 1e-13 and abs(oo - orr) <= 1e-13 * max(abs(orr), 1) and rel(ho, hr) < 1e-13
        so, po = O.sorted_mm(m[lo:hi]); sr, pr = R.sorted_mm(m[lo:hi])
        assert np.array_equal(so, sr)
        assert np.array_equal(m[lo:hi][po], m[lo:hi][pr])


@needs_ref
@pytest.mark.parametrize("solver", [1, 2])
def test_oracle_vs_live_reference_training(solver):
    ds = dataset("ragged")
    k, lam = 8, 15.0
    U = ob.ref_initial(ds.d1, k); V = ob.ref_initial(ds.d2, k)
    X, XT = to_csr(ds.train), to_csr(ds.test)
    a = ob.oracle().train(solver, X, XT, U, V, lam, 3)
    b = ob.reference().train(solver, X, XT, U, V, lam, 3)
    assert np.all(np.abs(a["obj"] - b["obj"]) <= 1e-12 * np.abs(b["obj"]))
    assert rel(a["U"], b["U"]) < 1e-12 and rel(a["V"], b["V"]) < 1e-12
    assert np.allclose(a["evals"], b["evals"], rtol=0, atol=1e-12)


@needs_ref
@pytest.mark.parametrize("solver", [1, 2])
@pytest.mark.parametrize("stepsize", [48.0, 1.0e6])
def test_oracle_vs_live_reference_line_search_branches(stepsize, solver):
    """The oracle's rarely taken branches against the REAL reference objects: parameter.stepsize = 48 makes both line
    searches halve several times, 1e6 makes update_V(_new) reject all 20 trials in the first iteration (V kept, stale
    scores handed to update_U, pcrpp.cpp:443) and every user keep its 20th trial (:814).  Identical factors after three
    iterations mean identical branch decisions.  (tests/test_gpu_parity.py::test_line_search_branches runs the same
    cases on the device against the oracle.)"""
    ds = dataset("tiny")
    k, lam = 7, 50.0
    U = ob.ref_initial(ds.d1, k); V = ob.ref_initial(ds.d2, k)
    X = to_csr(ds.train)
    a = ob.oracle().train(solver, X, None, U, V, lam, 3, do_predict=0, stepsize=stepsize)
    b = ob.reference().train(solver, X, None, U, V, lam, 3, do_predict=0, stepsize=stepsize)
    assert np.all(np.abs(a["obj"] - b["obj"]) <= 1e-12 * np.abs(b["obj"]))
    assert rel(a["U"], b["U"]) < 1e-12 and rel(a["V"], b["V"]) < 1e-12
    cnt = a["counters"]
    assert all(int(c[1]) > 1 for c in cnt)                                # V line search really took several trials
    assert (int(cnt[0][2]) == 0) == (stepsize > 1e3)                      # ... and was rejected outright with 1e6


def test_eval_pair_counts_match_brute_force():
    """The per-user integer pair-error counts the scale fixtures and the GPU tests rely on (orc_eval's `per_user`), against a
    literal numpy restatement of util.cpp:467-479 -- including exact score ties, which count as errors."""
    from tests.util import np_init, to_csr, dataset
    ds = dataset("tiny")
    U, V = np_init(ds.d1, ds.d2, 4, seed=1, scale=0.6)
    V[5] = V[7]; U[3] = 0.0                                  # exact ties
    out, counts, per_user = ob.oracle().eval(to_csr(ds.train), U, V, 10, want_counts=True)
    rp = ds.train.row_ptr
    tot = 0.0; users = 0
    for u in range(ds.d1):
        a, b = int(rp[u]), int(rp[u + 1])
        s = (U[u] * V[ds.train.item[a:b]]).sum(1) if b > a else np.zeros(0)
        # the oracle accumulates the dot product in the reference's loop order; recompute it the same way for exact ties
        s = np.array([sum(U[u, t] * V[ds.train.item[e], t] for t in range(U.shape[1])) for e in range(a, b)])
        v = ds.train.rating[a:b]
        err = 0
        for j in range(b - a):
            for q in range(j + 1, b - a):
                if (s[j] >= s[q] and v[j] < v[q]) or (s[j] <= s[q] and v[j] > v[q]):
                    err += 1
        assert err == per_user[u], u
        n = b - a
        if n * (n - 1) // 2 > 0:
            tot += err / (n * (n - 1) / 2); users += 1
    assert abs(out[0] - tot / users) < 1e-12


@pytest.mark.parametrize("case", ["netflix005", "powerlaw001", "ml1m_pcr"])
def test_scale_fixtures_are_self_consistent(case):
    """tests/golden/scale_*.npz (read by the GPU scale tests): the single-threaded restatement and the unmodified reference
    (race-free objects, all threads) must tell the same story inside the fixture, and the committed generator must still
    reproduce the data set the fixture was computed on."""
    import hashlib
    import os
    from primalcr_b200.data import synth_dataset
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "scale_%s.npz" % case))
    iters = int(g["iters"])
    assert g["obj"].shape == (iters + 1,) and g["counters"].shape == (iters, 8) and g["evals"].shape == (iters + 1, 4)
    spread = np.abs(g["obj"] - g["ref_obj"]) / np.abs(g["obj"])
    assert spread[0] < 1e-12 and spread[1] < 1e-10          # iteration 0 / 1: the reference reproduces itself
    assert spread.max() < (1e-5 if case == "powerlaw001" else 1e-12)
    assert np.all(np.diff(g["obj"]) < 0)                     # the objective decreases
    assert np.all(g["counters"][:, 0] <= 10) and np.all(g["counters"][:, 1] <= 20)
    d1 = int(g["d1"])
    assert g["err_train"].shape == (iters + 1, d1) and g["collapsed"].shape[0] == iters + 1
    collapsed = np.unpackbits(g["collapsed"], axis=1)[:, :d1]
    assert collapsed[0].sum() == 0                           # nobody has collapsed at the N(0,1) init
    if "ref1_obj" in g:                                       # the unmodified reference on one thread == the restatement, bitwise
        assert np.array_equal(g["ref1_obj"][1:], g["obj"][1:]) and np.array_equal(g["ref1_evals"], g["evals"])
    ds = synth_dataset(str(g["shape"]), scale=float(g["scale"]), device="cpu")
    h = hashlib.sha1()
    for a in (ds.train.row_ptr, ds.train.item, ds.train.rating, ds.test.row_ptr, ds.test.item, ds.test.rating):
        h.update(np.ascontiguousarray(a).tobytes())
    assert h.hexdigest() == str(g["digest"])
    # the evaluation means stored in the fixture follow from its own per-user integers
    ln = ds.train.lens().astype(np.float64); pairs = ln * (ln - 1) / 2
    for i in range(iters + 1):
        mean = (g["err_train"][i][pairs > 0] / pairs[pairs > 0]).mean()
        assert abs(mean - g["evals"][i, 0]) < 1e-12
