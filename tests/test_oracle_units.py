"""Oracle unit checks (CPU): the small routines against closed forms.
PARITY UNPINNED: the reference ships no goldens; these pin the restatement to analytic answers."""
import ctypes as C
import math

import numpy as np
import pytest

from oracle.oracle_py import load


@pytest.fixture(scope="module")
def lib():
    L = load()
    L.orc_qdr_src_2.restype = C.c_double
    L.orc_qdr_src_2.argtypes = [C.c_double] * 6
    L.orc_bplanck.restype = C.c_double
    L.orc_bplanck.argtypes = [C.c_double] * 2
    L.orc_hunt.restype = None
    L.orc_hunt.argtypes = [C.POINTER(C.c_double), C.c_int, C.c_double, C.POINTER(C.c_int)]
    return L


def test_qdr_src_2_constant_source(lib):
    # transfer.F:1498: constant S and alpha => I = I0 e^-tau + S (1 - e^-tau)
    for tau in (1e-3, 0.1, 1.0, 30.0):
        alp, S, ds, I0 = tau, 2.5, 1.0, 0.7
        got = lib.orc_qdr_src_2(I0, S * alp, alp, S * alp, alp, ds)
        want = I0 * math.exp(-tau) + S * (1.0 - math.exp(-tau))
        assert abs(got - want) <= 1e-14 * want


def test_qdr_src_2_thin_limits(lib):
    # dtau <= 1e-6: a=b=dtau/2, xp=1-dtau ; dtau <= 1e-9 (REAL literal): Q = 0.5 (j1+j2) ds
    j1, j2, ds = 3.0, 5.0, 2.0
    got = lib.orc_qdr_src_2(1.0, j1, 1e-12, j2, 1e-12, ds)
    assert got == 1.0 * (1.0 - 2e-12) + 0.5 * (j1 + j2) * ds
    a = 2e-7
    got = lib.orc_qdr_src_2(0.0, j1, a, j2, a, ds)
    dtau = a * ds
    assert got == min(0.5 * dtau * (j1 / a) + 0.5 * dtau * (j2 / a), 0.5 * (j1 + j2) * ds)


def test_qdr_src_2_bracket_and_zero_opacity(lib):
    # alpha = 0 everywhere: pure emission; bracket Q <= 0.5 (j1+j2) ds
    assert lib.orc_qdr_src_2(2.0, 1.0, 0.0, 3.0, 0.0, 4.0) == 2.0 + 8.0
    # maser (negative alpha) takes the thin branch with xp = 1 - dtau > 1
    got = lib.orc_qdr_src_2(1.0, 0.0, -0.1, 0.0, -0.1, 1.0)
    assert got == 1.0 * (1.0 + 0.1) + 0.0


def test_bplanck_literals(lib):
    # setup.F:937-952
    assert lib.orc_bplanck(0.0, 1e13) == 0.0
    nu, T = 6.3e13, 500.0
    want = 1.47455e-47 * nu * nu * nu / (math.exp(4.7989e-11 * nu / T) - 1.0) + 1e-290
    assert lib.orc_bplanck(T, nu) == want


def test_hunt_brackets(lib):
    xx = np.array([1.0, 2.0, 4.0, 8.0, 16.0])
    p = xx.ctypes.data_as(C.POINTER(C.c_double))
    for guess in (0, 1, 3, 5, 99):
        for x, want in ((0.5, 0), (1.5, 1), (3.0, 2), (9.0, 4), (20.0, 5)):
            j = C.c_int(guess)
            lib.orc_hunt(p, 5, x, C.byref(j))
            assert j.value == want
    # tie behaviour (nrecip.F:157): bisection uses '>', hunting uses '>='
    j = C.c_int(0)
    lib.orc_hunt(p, 5, 4.0, C.byref(j))
    assert j.value == 2
    j = C.c_int(3)
    lib.orc_hunt(p, 5, 4.0, C.byref(j))
    assert j.value == 3
