"""Shared helpers for the test-suite: small deterministic data sets and oracle adapters."""
import numpy as np

from oracle import bindings as ob
from primalcr_b200.data import Dataset, Ratings, synth_dataset


def to_csr(R: Ratings) -> ob.Csr:
    return ob.Csr(R.d1, R.d2, R.row_ptr, R.item.astype(np.int64), R.rating)


def init_factors(d1, d2, k, scale=1.0):
    """The reference's own init stream (util.cpp:80-93) via the product's host helper."""
    from primalcr_b200 import api
    U = api.reference_init(d1, k) * scale
    V = api.reference_init(d2, k) * scale
    return U, V


def np_init(d1, d2, k, seed=0, scale=1.0):
    rng = np.random.default_rng(seed)
    return rng.standard_normal((d1, k)) * scale, rng.standard_normal((d2, k)) * scale


def ragged_dataset(seed=3, d2=6000, real_valued=False) -> Dataset:
    """Edge cases the reference meets: empty users, 1-rating users, users whose ratings are all equal,
    a class-L user (1024 < len <= 4096) and a heavy user (len > 4096), exact score ties come from the tests."""
    rng = np.random.default_rng(seed)
    lens = [0, 1, 2, 3, 0, 7, 31, 32, 33, 64, 100, 255, 256, 257, 1023, 1024, 1025, 1500, 4096, 4097, 5000, 0, 5, 1]
    users, items, vals = [], [], []
    for u, n in enumerate(lens):
        it = np.sort(rng.choice(d2, size=n, replace=False))
        if real_valued:
            v = np.round(rng.standard_normal(n) * 1.7, 3)
        else:
            v = rng.integers(1, 6, size=n).astype(np.float64)
        if u in (5, 22):
            v[:] = 4.0            # all ratings equal: no comparable pair (PCR skips the user, pcr.cpp:552)
        users.append(np.full(n, u)); items.append(it); vals.append(v)
    d1 = len(lens)
    train = Ratings.from_coo(d1, d2, np.concatenate(users), np.concatenate(items), np.concatenate(vals))
    # test set: up to 12 ratings per user, file order (unsorted items) inside a user
    tu, ti, tv = [], [], []
    for u in range(d1):
        n = int(rng.integers(0, 13))
        a, b = int(train.row_ptr[u]), int(train.row_ptr[u + 1])
        if b > a and len(np.unique(train.rating[a:b])) < 2:
            n = 0     # no comparable pair => the Newton step sends u_i to ~0: test scores would be rounding noise
        it = rng.choice(d2, size=n, replace=False)
        tu.append(np.full(n, u)); ti.append(it)
        tv.append(rng.integers(1, 6, size=n).astype(np.float64) if not real_valued else np.round(rng.standard_normal(n), 2))
    from primalcr_b200.data import csr_in_file_order
    test = csr_in_file_order(d1, d2, np.concatenate(tu), np.concatenate(ti), np.concatenate(tv))
    return Dataset(train, test, name="ragged")


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    den = np.max(np.abs(b)) if b.size else 1.0
    return float(np.max(np.abs(a - b)) / (den if den > 0 else 1.0)) if a.size else 0.0


_cache = {}


def dataset(name):
    if name not in _cache:
        if name == "ragged":
            _cache[name] = ragged_dataset()
        elif name == "ragged_real":
            _cache[name] = ragged_dataset(seed=5, real_valued=True)
        else:
            _cache[name] = synth_dataset(name)
    return _cache[name]
