"""GPU: the C ABI used from plain C (examples/c_abi_demo.c: no Python, no torch in the process that calls the library) --
outputs compared with the oracle here."""
import os
import subprocess

import numpy as np
import pytest

from _util import TOL
from oracle import layers_np as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_plain_c_caller(tmp_path):
    exe = str(tmp_path / "c_abi_demo")
    csrc = os.path.join(ROOT, "recalgorithm_b200", "csrc")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    subprocess.run(["gcc", "-O2", "-I", os.path.join(ROOT, "include"), "-I", f"{cuda}/include", os.path.join(ROOT, "examples", "c_abi_demo.c"),
                    "-o", exe, "-L", csrc, "-lctr_b200", "-L", f"{cuda}/lib64", "-lcudart", f"-Wl,-rpath,{csrc}", f"-Wl,-rpath,{cuda}/lib64"],
                   check=True, capture_output=True)
    rng = np.random.default_rng(12)
    B, F, D = 77, 9, 16
    rows = rng.integers(3, 20, size=F)
    off = np.concatenate([[0], np.cumsum(rows)]).astype(np.int64)
    table = rng.standard_normal((int(off[-1]), D)).astype(np.float32)
    ids = np.stack([rng.integers(-1, rows[f], size=B) for f in range(F)], 1).astype(np.int64)
    d_tile = rng.standard_normal((B, F, D)).astype(np.float32)
    d_fm2 = rng.standard_normal(B).astype(np.float32)
    with open(tmp_path / "in.bin", "wb") as f:
        f.write(np.array([B, F, D, off[-1]], np.int64).tobytes() + off.tobytes() + ids.tobytes() + table.tobytes() + d_tile.tobytes() + d_fm2.tobytes())
    proc = subprocess.run([exe, str(tmp_path / "in.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True, timeout=120)
    assert proc.returncode == 0, proc.stderr
    assert "kernels launched" in proc.stderr
    out = np.fromfile(tmp_path / "out.bin", dtype=np.float32)
    n = B * F * D
    tile, fm2, rg = out[:n].reshape(B, F, D), out[n:n + B], out[n + B:].reshape(B, F, D)
    e = O.embedding_lookup(table, ids, off)
    assert np.array_equal(tile, e)                                               # gathered rows bit-exact
    ref = O.fm2_fwd(e.astype(np.float64))[:, 0]
    assert np.abs(fm2 - ref).max() <= TOL * np.abs(ref).max()
    want = d_tile.astype(np.float64) + O.fm2_bwd(e.astype(np.float64), d_fm2.astype(np.float64))
    assert np.abs(rg - want).max() <= TOL * np.abs(want).max()
