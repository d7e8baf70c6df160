"""CPU runs of libtsl kernels through the CUDA-on-CPU shim tests/csrc/cuda_emu.h (one std::thread per CUDA thread): the kernels with
real thread cooperation -- pivoted LU panels, shared-memory tiles, warp shuffles -- are exercised here, where there is no GPU, with
the same launch sequences the library uses.  The GPU tests repeat the comparisons on the real device."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "csrc", "_build")


def _build(name):
    os.makedirs(BUILD, exist_ok=True)
    src = os.path.join(HERE, "csrc", name + ".cpp")
    out = os.path.join(BUILD, "lib" + name + ".so")
    deps = [src, os.path.join(HERE, "csrc", "cuda_emu.h")]
    csrc = os.path.join(os.path.dirname(HERE), "thinshelllab_b200", "csrc")
    deps += [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(".cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.check_call(["g++", "-O1", "-std=c++20", "-pthread", "-shared", "-fPIC", src, "-o", out])
    return C.CDLL(out)


@pytest.mark.parametrize("n", [1, 7, 32, 33, 70, 129])
def test_dense_lu_kernels_match_numpy(n):
    """blocked right-looking LU with partial pivoting (tsl_dense_kernels.cuh) + triangular solves against numpy.linalg.solve"""
    L = _build("emu_dense")
    rng = np.random.default_rng(n)
    A = rng.standard_normal((n, n))
    if n > 1:
        A[0, 0] = 0.0                                 # forces a row exchange in the first panel
    if n > 40:
        A[35, :36] = 0.0; A[35, 35] = 1e-9            # a tiny pivot candidate inside the second panel
    b = rng.standard_normal(n)
    x = np.zeros(n)
    Af = np.asfortranarray(A)
    info = L.emu_lu_solve(n, Af.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p), 128)
    assert info == 0
    ref = np.linalg.solve(A, b)
    assert np.abs(x - ref).max() <= 1e-9 * np.abs(ref).max() * max(1.0, np.linalg.cond(A) * 1e-6)


def test_dense_lu_flags_singular_matrix():
    L = _build("emu_dense")
    n = 40
    A = np.random.default_rng(0).standard_normal((n, n))
    A[:, 5] = 0.0
    b = np.ones(n); x = np.zeros(n)
    Af = np.asfortranarray(A)
    assert L.emu_lu_solve(n, Af.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p), 64) == 1
