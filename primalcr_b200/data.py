"""Ratings data in the reference's own formats, plus the deterministic synthetic generators of SURVEY.md 8(d).

Formats kept byte-compatible with the reference (wuliwei9278/primalCR):
  * ``data_dir/meta``          -- util.cpp:9-21: ``m n`` / ``nnz_train train_file`` / ``nnz_test test_file``
  * ratings files              -- util.h:118-132, 360-371: one ``user item rating`` per line, 1-based ids;
                                  the test file must be sorted by user (util.cpp:257-266)
  * model file                 -- util.cpp:30-51 via pmf-train.cpp:297-310:
                                  ``int64 d1, int64 k, d1*k f64 (U row-major), int64 d2, int64 k, d2*k f64 (V)``

Everything here is host-side plumbing (numpy / torch tensors); the hot path lives in csrc/.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np


@dataclass
class Ratings:
    """CSR by user (the layout of the reference's SparseMat, util.h:390-413): items ascending inside a user."""
    d1: int
    d2: int
    row_ptr: np.ndarray   # int64 [d1+1]
    item: np.ndarray      # int32 [nnz]
    rating: np.ndarray    # float64 [nnz]

    @property
    def nnz(self) -> int:
        return int(self.row_ptr[-1])

    def lens(self) -> np.ndarray:
        return np.diff(self.row_ptr)

    def users(self) -> np.ndarray:
        return np.repeat(np.arange(self.d1, dtype=np.int64), self.lens())

    def slice_users(self, u0: int, u1: int) -> "Ratings":
        """Users [u0, u1) as a stand-alone CSR (item ids stay global)."""
        a, b = int(self.row_ptr[u0]), int(self.row_ptr[u1])
        return Ratings(u1 - u0, self.d2, (self.row_ptr[u0:u1 + 1] - a).astype(np.int64),
                       np.ascontiguousarray(self.item[a:b]), np.ascontiguousarray(self.rating[a:b]))

    @staticmethod
    def empty(d1: int, d2: int) -> "Ratings":
        return Ratings(d1, d2, np.zeros(d1 + 1, np.int64), np.zeros(0, np.int32), np.zeros(0, np.float64))

    @staticmethod
    def from_coo(d1: int, d2: int, users, items, ratings) -> "Ratings":
        """Sort by (user, item) exactly like smat_t::load_from_iterator (util.h:240) and build the CSR."""
        users = np.asarray(users, np.int64); items = np.asarray(items, np.int64)
        ratings = np.asarray(ratings, np.float64)
        order = np.lexsort((items, users))
        users, items, ratings = users[order], items[order], ratings[order]
        row_ptr = np.zeros(d1 + 1, np.int64)
        np.cumsum(np.bincount(users, minlength=d1), out=row_ptr[1:])
        return Ratings(d1, d2, row_ptr, items.astype(np.int32), ratings)


@dataclass
class Dataset:
    train: Ratings
    test: Ratings
    name: str = "synthetic"

    @property
    def d1(self): return self.train.d1

    @property
    def d2(self): return self.train.d2


# ----------------------------------------------------------------------------- reference text format

def write_ratings_file(path: str, R: Ratings) -> None:
    users = R.users() + 1
    items = R.item.astype(np.int64) + 1
    vals = R.rating
    if np.all(vals == np.rint(vals)):          # integer ratings: pandas' C writer (10 M lines in a few seconds)
        import pandas as pd
        pd.DataFrame({"u": users, "i": items, "r": vals.astype(np.int64)}).to_csv(path, sep=" ", header=False, index=False)
        return
    txt = np.array(["%d %d %.17g" % t for t in zip(users, items, vals)])
    with open(path, "w") as f:
        f.write("\n".join(txt.tolist()))
        f.write("\n")


def write_reference_dir(path: str, ds: Dataset) -> None:
    """Writes ``meta``, ``training.ratings`` and (if any) ``test.ratings`` in the reference layout."""
    os.makedirs(path, exist_ok=True)
    write_ratings_file(os.path.join(path, "training.ratings"), ds.train)
    with open(os.path.join(path, "meta"), "w") as f:
        f.write("%d %d\n%d training.ratings\n" % (ds.d1, ds.d2, ds.train.nnz))
        if ds.test.nnz:
            f.write("%d test.ratings\n" % ds.test.nnz)
    if ds.test.nnz:
        write_ratings_file(os.path.join(path, "test.ratings"), ds.test)


def read_ratings_file(path: str, nnz: int | None = None):
    import pandas as pd
    df = pd.read_csv(path, sep=r"\s+", header=None, names=["u", "i", "r"], nrows=nnz,
                     dtype={"u": np.int64, "i": np.int64, "r": np.float64}, engine="c")
    return df["u"].to_numpy() - 1, df["i"].to_numpy() - 1, df["r"].to_numpy()


def read_reference_dir(path: str) -> Dataset:
    """Python mirror of load() util.cpp:6-25 (training entries are sorted by (user,item); the test set is
    taken in file order and must already be grouped by user, util.cpp:257-266)."""
    with open(os.path.join(path, "meta")) as f:
        tok = f.read().split()
    d1, d2 = int(tok[0]), int(tok[1])
    u, i, r = read_ratings_file(os.path.join(path, tok[3]), int(tok[2]))
    train = Ratings.from_coo(d1, d2, u, i, r)
    test = Ratings.empty(d1, d2)
    if len(tok) >= 6:
        u, i, r = read_ratings_file(os.path.join(path, tok[5]), int(tok[4]))
        test = csr_in_file_order(d1, d2, u, i, r)
    return Dataset(train, test, name=os.path.basename(os.path.normpath(path)))


def csr_in_file_order(d1, d2, users, items, ratings) -> Ratings:
    """convert(testset_t&) util.cpp:250-274: keeps the file order inside a user; users must be non-decreasing."""
    users = np.asarray(users, np.int64)
    if len(users) > 1 and np.any(np.diff(users) < 0):
        raise ValueError("test ratings must be sorted by user (reference util.cpp:257-266)")
    row_ptr = np.zeros(d1 + 1, np.int64)
    np.cumsum(np.bincount(users, minlength=d1), out=row_ptr[1:])
    return Ratings(d1, d2, row_ptr, np.asarray(items, np.int32), np.asarray(ratings, np.float64))


# ----------------------------------------------------------------------------- model file

def save_model(path: str, U: np.ndarray, V: np.ndarray) -> None:
    with open(path, "wb") as f:
        for M in (U, V):
            M = np.ascontiguousarray(M, np.float64)
            np.array(M.shape, np.int64).tofile(f)
            M.tofile(f)


def load_model(path: str):
    with open(path, "rb") as f:
        out = []
        for _ in range(2):
            m, n = np.fromfile(f, np.int64, 2)
            out.append(np.fromfile(f, np.float64, int(m * n)).reshape(int(m), int(n)))
    return out[0], out[1]


# ----------------------------------------------------------------------------- sharding

def shard_bounds(row_ptr: np.ndarray, world: int) -> np.ndarray:
    """Contiguous user ranges balanced by nnz: bounds[r]..bounds[r+1] is rank r's slice."""
    d1 = len(row_ptr) - 1
    nnz = int(row_ptr[-1])
    targets = (np.arange(1, world) * (nnz / world))
    cuts = np.searchsorted(row_ptr, targets, side="left")
    b = np.concatenate([[0], np.clip(cuts, 0, d1), [d1]]).astype(np.int64)
    return np.maximum.accumulate(b)


# ----------------------------------------------------------------------------- synthetic generator

SHAPES = {
    # name: d1, d2, nnz_train, sigma (lognormal degree spread), seed, degree law
    "tiny":       dict(d1=300, d2=120, nnz=6_000, sigma=1.0, seed=7, law="lognormal"),
    "ml1m":       dict(d1=6040, d2=3952, nnz=939_809, sigma=1.0, seed=12345, law="lognormal"),
    "netflix":    dict(d1=480_189, d2=17_770, nnz=100_000_000, sigma=1.05, seed=20170813, law="lognormal"),
    "yahoo":      dict(d1=1_000_000, d2=625_000, nnz=250_000_000, sigma=1.05, seed=20170814, law="lognormal"),
    "powerlaw":   dict(d1=2_000_000, d2=500_000, nnz=500_000_000, sigma=0.0, seed=20170815, law="pareto",
                       max_deg=100_000),
}


def synth_dataset(shape: str | dict, scale: float = 1.0, device: str = "cpu", test_per_user: int = 10,
                  rank: int = 16, levels: int = 5) -> Dataset:
    """Deterministic synthetic ratings of a named shape.

    Degrees: lognormal(sigma) (or Pareto, clipped at max_deg) rescaled so the TRAIN ratings sum to ~nnz,
    clipped to [1, d2]; items per user are drawn without replacement from a Zipf-like popularity
    (weight ~ rank^-0.8 under a random rank->id permutation); rating = clip(round(3 + 1.2*(u.v/sqrt(rank)
    + 0.5*eps)), 1, levels) from a rank-`rank` ground truth; `test_per_user` extra ratings per user with
    at least 2*test_per_user ratings are held out as the test set.  `scale` < 1 keeps d2 and shrinks the
    number of users and ratings proportionally (a user subsample of the same distribution).

    Runs on torch (CPU for tests, CUDA for the full-size bench); the stream depends on (seed, device type).
    """
    import torch
    cfg = dict(SHAPES[shape]) if isinstance(shape, str) else dict(shape)
    name = shape if isinstance(shape, str) else cfg.get("name", "custom")
    d1 = max(1, int(round(cfg["d1"] * scale))); d2 = int(cfg["d2"])
    nnz = max(1, int(round(cfg["nnz"] * scale)))
    dev = torch.device(device)
    g = torch.Generator(device=dev); g.manual_seed(int(cfg["seed"]))
    f64 = torch.float64

    # ---- degrees (total = train + held-out)
    if cfg.get("law", "lognormal") == "pareto":
        u = torch.rand(d1, generator=g, device=dev, dtype=f64)
        raw = (1.0 - u).pow(-1.0 / 1.2)            # Pareto(alpha=1.2), minimum 1
    else:
        raw = torch.exp(cfg["sigma"] * torch.randn(d1, generator=g, device=dev, dtype=f64))
    max_deg = min(int(cfg.get("max_deg", d2)), d2)
    target_total = nnz + test_per_user * d1
    fac = target_total / float(raw.sum())
    for _ in range(6):                              # clipping changes the sum: rescale a few times
        deg = torch.clamp(torch.round(raw * fac), 1, max_deg)
        fac *= target_total / float(deg.sum())
    deg = deg.to(torch.int64)

    # ---- items without replacement under Zipf-like popularity
    ranks = torch.arange(1, d2 + 1, device=dev, dtype=f64)
    w = ranks.pow(-0.8)
    perm = torch.randperm(d2, generator=g, device=dev)
    # sequential prefix sum on the host: a multi-tile CUDA scan (d2 > a few thousand) associates its fp64 partial sums in a
    # timing-dependent order, so the cdf -- and a handful of the items drawn from it -- differed between two GPU boxes
    # (Yahoo shape, round 2: objective at iteration 0 equal only to 3e-8).  Bit-identical to torch's CPU cumsum.
    cdf = torch.from_numpy(np.cumsum((w / w.sum()).cpu().numpy())).to(dev)
    heavy = deg > d2 // 4
    keys = torch.empty(0, dtype=torch.int64, device=dev)
    # heavy users: exact weighted sampling without replacement (exponential keys, smallest first)
    hidx = torch.nonzero(heavy).flatten()
    for c0 in range(0, len(hidx), 256):
        hu = hidx[c0:c0 + 256]
        e = -torch.log(torch.rand(len(hu), d2, generator=g, device=dev, dtype=f64).clamp_min(1e-300)) / w
        order = torch.argsort(e, dim=1)
        take = torch.arange(d2, device=dev)[None, :] < deg[hu][:, None]
        it = perm[order][take]
        us = hu[:, None].expand(-1, d2)[take]
        keys = torch.cat([keys, us * d2 + it])
    # light users: rejection rounds (draw with replacement, de-duplicate, redraw the deficit)
    have = torch.zeros(d1, dtype=torch.int64, device=dev)
    want = torch.where(heavy, torch.zeros_like(deg), deg)
    for rnd in range(12):
        deficit = torch.clamp(want - have, min=0)
        if int(deficit.sum()) == 0:
            break
        draw = torch.where(deficit > 0, deficit + deficit // 2 + 4, deficit)
        us = torch.repeat_interleave(torch.arange(d1, device=dev), draw)
        r = torch.rand(len(us), generator=g, device=dev, dtype=f64)
        it = perm[torch.searchsorted(cdf, r).clamp_max(d2 - 1)]
        lk = torch.unique(torch.cat([keys, us * d2 + it]))
        keys = lk
        cnt = torch.bincount(keys // d2, minlength=d1)
        have = torch.where(heavy, torch.zeros_like(cnt), cnt)
    # trim users that overshot: keep a random subset of size deg
    us = keys // d2
    pri = torch.rand(len(keys), generator=g, device=dev, dtype=f64)
    order = torch.argsort(us.to(f64) + pri * 0.999999)    # user-major, random inside a user
    us_o = us[order]
    start = torch.zeros(d1 + 1, dtype=torch.int64, device=dev)
    start[1:] = torch.cumsum(torch.bincount(us, minlength=d1), 0)
    rank_in_user = torch.arange(len(keys), device=dev) - start[us_o]
    keep = rank_in_user < deg[us_o]
    # held-out split: the first `test_per_user` of the random order, for users with enough ratings
    is_test = keep & (rank_in_user < test_per_user) & (deg[us_o] >= 2 * test_per_user)
    sel_keys = keys[order]

    gt_u = torch.randn(d1, rank, generator=g, device=dev, dtype=torch.float32)
    gt_v = torch.randn(d2, rank, generator=g, device=dev, dtype=torch.float32)

    def ratings_for(k):
        uu, ii = k // d2, k % d2
        out = torch.empty(len(k), dtype=f64, device=dev)
        for c0 in range(0, len(k), 1 << 23):
            sl = slice(c0, c0 + (1 << 23))
            s = (gt_u[uu[sl]] * gt_v[ii[sl]]).sum(1).to(f64) / (rank ** 0.5)
            eps = torch.randn(len(s), generator=g, device=dev, dtype=f64)
            out[sl] = torch.clamp(torch.round(3.0 + 1.2 * (s + 0.5 * eps)), 1, levels)
        return uu, ii, out

    def to_csr(k):
        k, _ = torch.sort(k)
        uu, ii, rr = ratings_for(k)
        row_ptr = torch.zeros(d1 + 1, dtype=torch.int64, device=dev)
        row_ptr[1:] = torch.cumsum(torch.bincount(uu, minlength=d1), 0)
        return Ratings(d1, d2, row_ptr.cpu().numpy(), ii.to(torch.int32).cpu().numpy(), rr.cpu().numpy())

    train = to_csr(sel_keys[keep & ~is_test])
    test = to_csr(sel_keys[is_test]) if test_per_user > 0 else Ratings.empty(d1, d2)
    return Dataset(train, test, name="%s-shape synthetic (scale %g)" % (name, scale))
