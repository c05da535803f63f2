"""The C ABI from compiled C (examples/c_client.c): the header is valid C99, the library links without Python, fails
loudly without a device, and on the GPU a plain-C caller gets bit for bit what the Python mirror gets."""
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "mima_b200", "lib")
DATA = os.path.join(ROOT, "mima_b200", "data")


def _build(tmp_path):
    import __graft_entry__ as ge
    ge.build()
    exe = str(tmp_path / "c_client")
    cmd = ["gcc", "-std=c99", "-O2", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "c_client.c"), "-L", LIBDIR, "-lrrtmg_b200", f"-Wl,-rpath,{LIBDIR}", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


@pytest.mark.skipif(shutil.which("gcc") is None, reason="needs gcc")
def test_c_client_builds_and_fails_loudly_without_a_device(tmp_path):
    exe = _build(tmp_path)
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present; the run is covered by the gpu test")
    r = subprocess.run([exe, DATA, "none", "none"], capture_output=True, text=True)
    assert r.returncode == 1 and "failed with 5" in r.stderr        # RRTMG_B200_ERR_CUDA, no CPU fallback


@pytest.mark.gpu
def test_c_client_matches_the_python_mirror(gpu, tmp_path):
    from mima_b200.columns import make_columns
    exe = _build(tmp_path)
    c = make_columns("T42L40", nlon=48, nlat=3, night=True)
    inp, outp = tmp_path / "columns.bin", tmp_path / "fluxes.bin"
    with open(inp, "wb") as f:
        np.array([c.ncol, c.nlay], dtype=np.int32).tofile(f)
        for a in (c.play, c.plev, c.tlay, c.tlev, c.tsfc, c.h2o, c.o3, c.co2, c.albedo, c.coszen):
            np.asfortranarray(a, dtype=np.float64).ravel(order="F").tofile(f)
        np.array([c.adjes, c.scon], dtype=np.float64).tofile(f)
    r = subprocess.run([exe, DATA, str(inp), str(outp), repr(float(gpu.CP_AIR))], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "c_client: 144 columns x 40 layers" in r.stdout
    raw = np.fromfile(outp, dtype=np.float64)
    L, V = c.ncol * c.nlay, c.ncol * (c.nlay + 1)
    ref = list(gpu.sw_from_columns(c)) + list(gpu.lw_from_columns(c))
    pos = 0
    for i, want in enumerate(ref):
        n = L if i % 3 == 2 else V
        got = raw[pos:pos + n].reshape(want.shape, order="F")
        pos += n
        assert np.array_equal(got, want), i
    assert pos == raw.size
