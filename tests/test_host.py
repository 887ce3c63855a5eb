"""CPU tests of the host-side logic: file formats, model layout, sharding, the C-ABI library's exports."""
import os
import re
import subprocess

import numpy as np
import pytest

from oracle import bindings as ob
from primalcr_b200 import api
from primalcr_b200.data import (Dataset, Ratings, load_model, read_reference_dir, save_model, shard_bounds,
                                synth_dataset, write_reference_dir)
from tests.util import dataset

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    txt = open(os.path.join(ROOT, "include", "primalcr.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(primalcr_[a-z_A-Z0-9]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    L = api.lib()
    names = header_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), n
    assert sorted(api.SYMBOLS) == names
    assert b"sm_100a" in L.primalcr_version()


def test_compute_entry_points_fail_loudly_without_gpu(have_gpu):
    if have_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(api.PrimalCRError, match="no CPU fallback|CUDA"):
        api.Engine(api.Parameter(k=4))


def test_reference_init_matches_golden_stream():
    g = np.load(os.path.join(ROOT, "tests", "golden", "golden_tiny.npz"))
    U = api.reference_init(int(g["d1"]), int(g["k"]))
    assert np.array_equal(U, g["U0"])          # same libstdc++ engine + distribution => same bits as the reference CLI
    if ob.reference() is not None:
        assert np.array_equal(api.reference_init(50, 3), ob.ref_initial(50, 3))


def test_reference_dir_roundtrip(tmp_path):
    ds = dataset("tiny")
    write_reference_dir(str(tmp_path / "d"), ds)
    back = read_reference_dir(str(tmp_path / "d"))
    for a, b in ((ds.train, back.train), (ds.test, back.test)):
        assert np.array_equal(a.row_ptr, b.row_ptr) and np.array_equal(a.item, b.item) and np.array_equal(a.rating, b.rating)
    meta = open(tmp_path / "d" / "meta").read().split()
    assert meta[:4] == [str(ds.d1), str(ds.d2), str(ds.train.nnz), "training.ratings"]


def test_unsorted_training_file_is_sorted_like_the_reference(tmp_path):
    ds = dataset("tiny")
    d = tmp_path / "d"; write_reference_dir(str(d), Dataset(ds.train, Ratings.empty(ds.d1, ds.d2)))
    lines = open(d / "training.ratings").read().splitlines()
    rng = np.random.default_rng(0); rng.shuffle(lines)
    open(d / "training.ratings", "w").write("\n".join(lines) + "\n")
    back = read_reference_dir(str(d))
    assert np.array_equal(back.train.item, ds.train.item) and np.array_equal(back.train.rating, ds.train.rating)


def test_model_file_layout(tmp_path):
    rng = np.random.default_rng(1)
    U, V = rng.standard_normal((7, 3)), rng.standard_normal((5, 3))
    p = str(tmp_path / "m.model")
    save_model(p, U, V)
    raw = open(p, "rb").read()
    assert len(raw) == 32 + 8 * 3 * (7 + 5)                       # util.cpp:30-51
    assert np.frombuffer(raw[:16], np.int64).tolist() == [7, 3]
    U2, V2 = load_model(p)
    assert np.array_equal(U, U2) and np.array_equal(V, V2)
    exe = ob.ref_cli("omp-pmf-predict")
    if exe:                                                        # the reference's own predictor reads our model
        t = tmp_path / "t.ratings"; t.write_text("1 1 3\n7 5 2\n3 2 5\n")
        subprocess.run([exe, str(t), p, str(tmp_path / "out")], check=True)
        got = np.loadtxt(tmp_path / "out")
        want = np.array([U[0] @ V[0], U[6] @ V[4], U[2] @ V[1]])
        assert np.allclose(got, want, atol=1e-6)


def test_shard_bounds_balance_and_cover():
    ds = synth_dataset("ml1m")
    rp = ds.train.row_ptr
    for world in (1, 2, 4, 8):
        b = shard_bounds(rp, world)
        assert b[0] == 0 and b[-1] == ds.d1 and np.all(np.diff(b) >= 0) and len(b) == world + 1
        nnz = np.diff(rp[b])
        assert nnz.sum() == ds.train.nnz
        assert nnz.max() <= ds.train.nnz / world + ds.train.lens().max()
    parts = [ds.train.slice_users(int(b[r]), int(b[r + 1])) for r in range(8)]
    assert sum(p.nnz for p in parts) == ds.train.nnz
    assert np.array_equal(np.concatenate([p.item for p in parts]), ds.train.item)


def test_synthetic_generator_is_deterministic_and_well_formed():
    a, b = synth_dataset("tiny"), synth_dataset("tiny")
    assert np.array_equal(a.train.item, b.train.item) and np.array_equal(a.train.rating, b.train.rating)
    R = a.train
    assert R.rating.min() >= 1 and R.rating.max() <= 5 and R.item.max() < R.d2
    for u in range(R.d1):
        it = R.item[R.row_ptr[u]:R.row_ptr[u + 1]]
        assert np.all(np.diff(it) > 0)                              # ascending, no duplicates (util.h:240)
    pl = synth_dataset("powerlaw", scale=0.0005)
    assert pl.train.lens().max() <= 100_000


def test_fast_loader_matches_python_reader(tmp_path):
    """host/loader.hpp (parallel mmap parser, counting sort by user) == the Python mirror of the reference loader,
    on a shuffled training file (the reference sorts by (user, item), util.h:240) and a grouped test file."""
    ds = dataset("tiny")
    d = tmp_path / "d"; write_reference_dir(str(d), ds)
    lines = open(d / "training.ratings").read().splitlines()
    np.random.default_rng(3).shuffle(lines)
    open(d / "training.ratings", "w").write("\n".join(lines) + "\n")
    a = api.load_dir(str(d), 4)
    b = read_reference_dir(str(d))
    for x, y in ((a.train, b.train), (a.test, b.test)):
        assert x.d1 == y.d1 and x.d2 == y.d2
        assert np.array_equal(x.row_ptr, y.row_ptr) and np.array_equal(x.item, y.item) and np.array_equal(x.rating, y.rating)
    assert np.array_equal(a.train.item, ds.train.item) and np.array_equal(a.train.rating, ds.train.rating)


def test_fast_loader_real_valued_and_errors(tmp_path):
    d = tmp_path / "d"; d.mkdir()
    (d / "meta").write_text("3 4\n5 tr.txt\n2 te.txt\n")
    (d / "tr.txt").write_text("3 1 -0.5\n1 4 2.25e0\n1 2 1e-3\n2 3 4\n3 4 0.0576852\n")
    (d / "te.txt").write_text("1 1 5\n3 2 1.5\n")
    a = api.load_dir(str(d), 2)
    assert a.train.row_ptr.tolist() == [0, 2, 3, 5] and a.train.item.tolist() == [1, 3, 2, 0, 3]
    assert a.train.rating.tolist() == [1e-3, 2.25, 4.0, -0.5, 0.0576852]
    assert a.test.row_ptr.tolist() == [0, 1, 1, 2] and a.test.rating.tolist() == [5.0, 1.5]
    with pytest.raises(api.PrimalCRError):
        api.load_dir(str(tmp_path / "missing"))
    (d / "tr.txt").write_text("9 1 1\n")
    (d / "meta").write_text("3 4\n1 tr.txt\n")
    with pytest.raises(api.PrimalCRError, match="out of range"):
        api.load_dir(str(d))


def test_fast_loader_decimal_parsing_is_correctly_rounded(tmp_path):
    """The loader's exact-decimal fast path (<= 15 significant digits) and its strtod fallback must both return the
    correctly rounded double, i.e. what Python's float() -- and the reference's sscanf("%lf") -- give."""
    rng = np.random.default_rng(11)
    texts = ["0", "-0", "5", "4.5", "0.1", "-0.3", "3.14159265358979", "123456789012345", "0.000000000000001",
             "99999.9999999999", "1e-3", "2.5E2", "-7.25e+1", "1234567890.1234567890123", "+3.5", "007.500", ".5", "5."]
    for _ in range(3000):
        nd_int = int(rng.integers(0, 9)); nd_frac = int(rng.integers(0, 12))
        s = "".join(str(int(c)) for c in rng.integers(0, 10, size=nd_int)) or "0"
        if nd_frac:
            s += "." + "".join(str(int(c)) for c in rng.integers(0, 10, size=nd_frac))
        if rng.random() < 0.3:
            s = "-" + s
        texts.append(s)
    d = tmp_path / "d"; d.mkdir()
    (d / "meta").write_text("1 %d\n%d tr.txt\n" % (len(texts), len(texts)))
    (d / "tr.txt").write_text("".join("1 %d %s\n" % (j + 1, t) for j, t in enumerate(texts)))
    a = api.load_dir(str(d), 3)
    want = np.array([float(t) for t in texts])
    assert a.train.item.tolist() == list(range(len(texts)))
    assert np.array_equal(a.train.rating.view(np.uint64), want.view(np.uint64))


def test_text_matrix_writer_is_byte_identical_to_the_reference_stream(tmp_path):
    """U.txt / V.txt (pmf-train.cpp:276-295): `ofstream << double` at the default precision is printf's %g; the parallel
    writer must produce the same bytes, row blocks in order, for more rows than one block and awkward values."""
    rng = np.random.default_rng(5)
    M = rng.standard_normal((9001, 7)) * 10.0 ** rng.integers(-12, 12, size=(9001, 7))
    M[0, :] = [0.0, -0.0, 1.0, -1.5, 1e-5, 123456789.0, 0.1]
    M[1, :3] = [1e300, 5e-324, 999999.5]
    p = str(tmp_path / "U.txt")
    api.write_text_matrix(p, M)
    want = "".join(" ".join("%g" % x for x in row) + "\n" for row in M)
    assert open(p).read() == want
    ref = ob.ref_cli("omp-pmf-train")
    if ref is not None:                      # and the reference CLI itself agrees on the format (tiny run, U.txt side file)
        ds = dataset("tiny")
        write_reference_dir(str(tmp_path / "d"), ds)
        subprocess.run([ref, "-s", "2", "-k", "3", "-t", "0", "-p", "0", "-n", "1", str(tmp_path / "d"), str(tmp_path / "m")],
                       cwd=tmp_path, check=True, capture_output=True)
        U, V = load_model(str(tmp_path / "m"))
        api.write_text_matrix(str(tmp_path / "U_ours.txt"), U)
        assert open(tmp_path / "U_ours.txt").read() == open(tmp_path / "U.txt").read()
