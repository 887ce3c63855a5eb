"""Multi-GPU parity (needs >= 2 GPUs; skipped otherwise): users sharded over 2 GPUs by the C++ host driver (one thread +
one engine per GPU, NCCL allreduce of the V-side sums) must reproduce the single-GPU / reference trajectory."""
import os
import subprocess

import numpy as np
import pytest

from primalcr_b200.data import Dataset, Ratings, load_model, write_reference_dir
from tests.util import rel

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("solver", [2, 1])
@pytest.mark.parametrize("gpus", [2, 4])
def test_cli_sharded_over_gpus_matches_reference_golden(tmp_path, solver, gpus):
    if _ngpu() < gpus:
        pytest.skip("needs %d GPUs" % gpus)
    exe = os.path.join(ROOT, "primalcr_b200", "bin", "primalcr-train")
    g = np.load(os.path.join(ROOT, "tests", "golden", "golden_tiny.npz"))
    d1, d2, k, lam, iters = int(g["d1"]), int(g["d2"]), int(g["k"]), float(g["lam"]), int(g["iters"])
    ds = Dataset(Ratings(d1, d2, g["row_ptr"], g["item"], g["rating"]), Ratings(d1, d2, g["t_row_ptr"], g["t_item"], g["t_rating"]))
    write_reference_dir(str(tmp_path / "data"), ds)
    env = dict(os.environ, PRIMALCR_GPUS=str(gpus))
    out = subprocess.run([exe, "-s", str(solver), "-k", str(k), "-l", str(lam), "-t", str(iters), "-p", "1",
                          str(tmp_path / "data"), str(tmp_path / "model")], cwd=tmp_path, capture_output=True, text=True,
                         env=env, timeout=300)
    assert out.returncode == 0, out.stderr
    objs = [float(l.split()[-1]) for l in out.stdout.splitlines() if l.startswith("Iter ")]
    want = g["s%d_obj" % solver]
    assert len(objs) == iters + 1
    assert np.all(np.abs(np.array(objs) - want) <= 2e-5 * np.abs(want))
    te = [float(l.split()[-1]) for l in out.stdout.splitlines() if l.startswith("(Testing)")]
    assert np.all(np.abs(np.array(te) - g["s%d_evals" % solver][:, 3]) < 1e-4)
    U, V = load_model(str(tmp_path / "model"))
    assert rel(U, g["s%d_U" % solver]) < 1e-7 and rel(V, g["s%d_V" % solver]) < 1e-7
