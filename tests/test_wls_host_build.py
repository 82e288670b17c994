"""The product's FP64 active-set routine (dronesim_b200/csrc/ds_wls.cuh) compiled for the HOST with gcc and run
against the reference-generated fixture: iteration count and final working set integer-exact on every regular run.
(The same comparison runs on the GPU through ds_debug_wls; this one needs no device.)"""
import ctypes as C
import os
import shutil
import subprocess
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

HARNESS = r"""
#include <math.h>
#define __device__
#define __noinline__
#define __restrict__
struct DsWlsDev { double B[36]; double Wv[6]; double gamma; double pmin[6], pmax[6]; int n_u, n_v; };
#define DS_WLS_HOST_BUILD 1
#include "ds_wls.cuh"
extern "C" int wls_host(const double* B, const double* Wv, double gamma, const double* v, const double* cmd, double* u, int* W) {
  DsWlsDev P;
  for (int i = 0; i < 36; ++i) P.B[i] = B[i];
  for (int i = 0; i < 6; ++i) { P.Wv[i] = Wv[i]; P.pmin[i] = 0.0; P.pmax[i] = 1.0; }
  P.gamma = gamma; P.n_u = 6; P.n_v = 6;
  double umin[6], umax[6];
  for (int i = 0; i < 6; ++i) { umin[i] = 0.0 - cmd[i]; umax[i] = 1.0 - cmd[i]; u[i] = 0.0; }
  return ds_wls_alloc(&P, v, umin, umax, u, W);
}
"""


@pytest.mark.skipif(shutil.which("g++") is None, reason="no host compiler")
def test_product_wls_routine_on_host_matches_reference_fixture():
    g = dict(np.load(os.path.join(GOLD, "wls_cases.npz")))  # an NpzFile would re-read the archive on every access
    with tempfile.TemporaryDirectory() as td:
        src = open(os.path.join(ROOT, "dronesim_b200", "csrc", "ds_wls.cuh")).read()
        # the device header is replaced by the one struct the routine needs (declared in the harness)
        src = src.replace('#include "ds_device.cuh"', "")
        open(os.path.join(td, "ds_wls.cuh"), "w").write(src)
        open(os.path.join(td, "harness.cpp"), "w").write(HARNESS)
        so = os.path.join(td, "libwls_host.so")
        subprocess.run(["g++", "-O1", "-ffp-contract=off", "-shared", "-fPIC", "-o", so, os.path.join(td, "harness.cpp")],
                       check=True, cwd=td)
        lib = C.CDLL(so)
        lib.wls_host.restype = C.c_int
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))  # noqa: E731
        B, Wv = np.ascontiguousarray(g["rnd_B"]), np.ascontiguousarray(g["rnd_Wv"])
        reg = ~g["rnd_stale"]
        n_multi = 0
        for k in np.flatnonzero(reg):
            v = g["rnd_v"][k].astype(np.float64).copy()
            c = g["rnd_cmd"][k].astype(np.float64).copy()
            u, W = np.zeros(6), np.zeros(6, dtype=np.int32)
            it = lib.wls_host(dp(B), dp(Wv), C.c_double(100000.0), dp(v), dp(c), dp(u), W.ctypes.data_as(C.POINTER(C.c_int)))
            assert it == g["rnd_iter"][k], "case %d: %d iterations, reference %d" % (k, it, g["rnd_iter"][k])
            assert (W == g["rnd_W"][k]).all(), "case %d: working set differs" % k
            assert np.abs(u - g["rnd_du"][k]).max() <= 1e-9 * max(1.0, np.abs(g["rnd_du"][k]).max())
            n_multi += int(it > 1)
        assert reg.sum() >= 10000 and n_multi >= 1000
