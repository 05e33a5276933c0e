"""The translator behind the reference pin (oracle/ref/f90cxx.py) against the Fortran rules themselves: a test program
with one routine per language rule the Noah-MP sources rely on (tests/golden/f90cxx_semantics.F90, written for this
test) is translated, compiled with g++ and every result compared with the value the Fortran standard — or gfortran,
where the standard leaves the choice — prescribes.  Needs only g++."""
import math
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
f32 = np.float32


@pytest.fixture(scope="module")
def M(tmp_path_factory):
    d = tmp_path_factory.mktemp("f90cxx")
    cpp, so = str(d / "sem.cpp"), str(d / "libsem.so")
    subprocess.check_call([sys.executable, os.path.join(ROOT, "oracle", "ref", "f90cxx.py"), cpp,
                           os.path.join(ROOT, "tests", "golden", "f90cxx_semantics.F90")], stdout=subprocess.DEVNULL)
    subprocess.check_call(["g++", "-std=gnu++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared",
                           "-I" + os.path.join(ROOT, "oracle", "ref"), cpp, os.path.join(ROOT, "oracle", "ref", "ref_shim.cpp"),
                           "-o", so])
    from oracle.ref import refmodel
    return refmodel.RefModel(so)


def out(n):
    return np.zeros(n, f32)


def test_operator_precedence_and_association(M):
    o = out(8)
    a, b, c = f32(3.0), f32(0.7), f32(1.3)
    M.call("PRECEDENCE", a, b, c, o)
    assert o[0] == -(a * a)
    assert o[1] == (a - b) + c and o[2] == (a / b) * c
    assert o[3] == 512.0
    assert o[4] == -(a * b)
    assert o[5] == a + b * (c * c)
    assert o[6] == (a + b) * c
    assert o[7] == f32(1.0) / (a * a)


def test_integer_arithmetic_and_conversions(M):
    o = out(8)
    M.call("INTEGER_RULES", 7, 2, o)
    assert list(o) == [3.0, -3.0, 0.0, 3.5, -1.0, 0.0, -2.0, 2.0]


def test_declared_lower_bounds_sections_and_reductions(M):
    z = np.array([10, 20, 30, 40, 50, 60, 170], f32)   # Z(-2:4)
    o = out(6)
    M.call("LOWER_BOUNDS", 3, 4, -2, z, o)              # ISNOW = -2: the section is W(-1:0) = Z(-1:0) * 2
    assert o[0] == 7.0 and o[1] == 60.0                 # W(-2) untouched, W(0) = Z(0) * 2
    assert list(z) == [8, 19, 30, 41, 52, 63, 174]
    assert o[2] == 8.0 and o[3] == 174.0 and o[4] == 174.0
    assert o[5] == 1.0                                  # some soil layers above 100, some below


def test_do_loops(M):
    o = out(8)
    M.call("DO_LOOPS", 5, o)
    assert list(o) == [5.0, 6.0, 3.0, -1.0, 0.0, 4.0, 3.0, 6.0]


def test_arguments_by_reference(M):
    o = out(6)
    M.var("SEM_GLOBALS.SHARED_VEC")[:] = 0
    res = M.call("BY_REFERENCE", 2.0, o)
    # BUMP(Y=1, 2*X+1 = 5, V(0), V(1:2)): Y = 6, V(0) = 5, V(1) = 6, V(2) = -6
    assert list(o[:4]) == [6.0, 5.0, 6.0, -6.0]
    # BUMP(X, 0.5, SHARED_SCALAR, SHARED_VEC): X = 2.5 comes back through the dummy; the module variables are set
    assert res[0] == 2.5
    assert o[4] == 0.5 and o[5] == -2.5
    assert M.var("SEM_GLOBALS.SHARED_SCALAR")[0] == 0.5 and list(M.var("SEM_GLOBALS.SHARED_VEC")) == [2.5, -2.5]


def test_statement_function_and_use_association_hides_the_host(M):
    o = out(4)
    M.call("STATEMENT_FUNCTION_AND_HIDING", 280.0, o)
    assert o[0] == f32(280.0) - f32(273.16) and o[1] == 50.0
    assert o[2] == 2106.0                               # SEM_CONSTANTS' CICE, USEd inside the routine
    assert o[3] == f32(1.0) / f32(3.0)
    h = out(1)
    M.call("HOST_CONSTANT", h)
    assert h[0] == f32(2.094e6)                         # the host module's CICE elsewhere


def test_save_and_data(M):
    o = out(4)
    M.call("SAVED_AND_DATA", o)
    first = o[0]
    M.call("SAVED_AND_DATA", o)
    assert o[0] == first + 1                            # SAVE
    assert o[1] == f32(0.1) and o[2] == 20.0 and o[3] == 3.0


def test_where_elsewhere(M):
    a = np.asfortranarray(np.array([[0.0, 2.0], [-1.0, 1.4], [1.6, -3.0]], f32))   # A(3,2)
    l = np.asfortranarray(np.zeros((3, 2), np.int32))
    M.call("MASKS", 3, a, l)
    assert l.tolist() == [[1, -1], [-1, 1], [-1, 1]]


def test_min_max_with_nan_follow_gfortran(M):
    o = out(6)
    M.call("MINMAX_NAN", float("nan"), o)
    assert o[0] == f32(1e-4) and o[1] == f32(1e-4) and o[2] == 2.0
    assert o[3] == 3.0 and o[4] == 2.0 and o[5] == -2.0


def test_powers(M):
    o = out(6)
    x = f32(1.2345678)
    M.call("POWERS", x, 5, o)
    x2 = x * x
    assert o[0] == x * (x2 * x2) and o[1] == o[0]       # constant and run-time exponent alike: square-and-multiply
    assert abs(float(o[2]) - float(x) ** 2) <= 1.2e-7 * float(x) ** 2   # libm powf, within an ulp of x*x
    assert o[3] == 32.0
    assert o[4] == f32(1.0) / (x * x2)
    assert o[5] == f32(math.sqrt(float(x))) or abs(float(o[5]) - math.sqrt(float(x))) < 1e-7


def test_optional_arguments(M):
    o = out(3)
    M.call("OPTIONAL_ARGS", 1.0, o, None, None)
    assert list(o) == [1.0, 0.0, 0.0]
    M.call("OPTIONAL_ARGS", 1.0, o, 2.0, np.array([5.0, 6.0], f32))
    assert list(o) == [1.0, 2.0, 6.0]


def test_internal_procedure_sees_the_host(M):
    o = out(2)
    M.call("WITH_INTERNAL", 2.0, o)
    assert list(o) == [6.0, 8.0]


def test_goto_and_labels(M):
    o = out(2)
    M.call("GOTOS", -0.1, o)
    assert list(o) == [0.0, 4.0]
    M.call("GOTOS", 0.3, o)
    assert list(o) == [1.0, 4.0]


def test_error_fatal_becomes_an_exception(M):
    M.call("FATAL", 2)
    with pytest.raises(RuntimeError, match="too many things"):
        M.call("FATAL", 5)


def test_three_dimensional_sections(M):
    # A(ims:ime, kms:kme, jms:jme) with ims=2, kms=1, jms=3; element (I=3, :, J=4)
    a = np.asfortranarray(np.arange(3 * 2 * 2, dtype=f32).reshape(3, 2, 2, order="F"))
    want_col = a[1, :, 1].copy()
    col = out(2)
    M.call("ARRAYS_3D", 2, 4, 1, 2, 3, 4, 3, 4, a, col)
    assert list(col) == list(want_col)
    assert a[1, 1, 1] == 1.0 and a[1, 0, 1] == want_col[1] / (want_col[0] + want_col[1])
    assert a[0, 0, 0] == 0.0 and a[2, 1, 1] == 11.0      # nothing else touched
