"""The host side above the C-ABI in the compiled language this image has: include/noahmp_b200_driver.hpp mirrors the
reference's operator interface — `noahmplsm`, `NOAHMP_INIT`, `WTABLE_mmf_noahmp` with their positional dummy lists — and
integration/example_driver.cpp is a complete host program on it.  Both are generated from the same lists the header and
the Fortran shim are checked against (tests/test_boundary.py).  CPU: they are current and compile / link; GPU: the
program's results equal, bit for bit, those of the same case through the Python mirror."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from noahmp_b200 import _capi, _lib, synthetic as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gen(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_cpp_driver.py"), *args], capture_output=True,
                          text=True, check=True).stdout


def test_generated_sources_are_current():
    assert _gen() == open(os.path.join(ROOT, "include", "noahmp_b200_driver.hpp")).read()
    assert _gen("example") == open(os.path.join(ROOT, "integration", "example_driver.cpp")).read()


def test_positional_lists_follow_the_reference_order():
    hpp = open(os.path.join(ROOT, "include", "noahmp_b200_driver.hpp")).read()
    for fn, spec in (("noahmplsm", _capi.ARGS_SPEC), ("NOAHMP_INIT", _capi.INIT_SPEC), ("WTABLE_mmf_noahmp", _capi.WT_SPEC)):
        m = re.search(r"inline void " + fn + r"\(\s*(.*?)\)\s*\{", hpp, re.S)
        names = [p.split()[-1].lower() for p in m.group(1).split(",")]
        assert names == [n for n, _ in spec], fn


def _build(tmp_path, built):
    exe = str(tmp_path / "example_driver")
    lib_dir = os.path.dirname(_lib.SO_PATH)
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "integration", "example_driver.cpp"), "-L" + lib_dir, "-lnoahmp_b200",
                    "-Wl,-rpath," + lib_dir, "-o", exe], check=True)
    return exe


def test_example_compiles_and_links(built, tmp_path):
    exe = _build(tmp_path, built)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr


@pytest.mark.gpu
def test_cpp_host_program_equals_python_mirror(built, tables_usgs, tables_usgs_struct, tmp_path):
    import noahmp_b200
    ni, nj, nsteps = 24, 16, 3
    exe = _build(tmp_path, built)
    tb = tmp_path / "tables.bin"
    tb.write_bytes(bytes(memoryview(tables_usgs_struct)))
    r = subprocess.run([exe, str(tb), str(ni), str(nj), str(nsteps)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = dict(l.split(" ", 1) for l in r.stdout.strip().splitlines())

    # the same case through the Python mirror
    f = np.float32
    ii, jj = np.meshgrid(np.arange(ni), np.arange(nj))
    water, glacier = jj == 1, jj == 3
    A = {n: np.zeros(_capi.array_shape(n, ni, nj), _capi.array_dtype(n)) for n in _capi.ARRAY_NAMES}
    A["ivgtyp"][...] = np.where(water, 16, np.where(glacier, 24, 2 + (ii + 3 * jj) % 14))
    A["isltyp"][...] = np.where(water, 14, np.where(glacier, 16, 1 + (2 * ii + jj) % 12))
    A["xland"][...] = np.where(water, 2.0, 1.0)
    A["tmn"][...] = np.where(glacier, f(260.0), f(283.0) + f(0.1) * (ii % 7).astype(f))
    A["xlatin"][...] = f(30.0) + jj.astype(f)
    A["vegfra"][...] = f(20.0) + ((5 * ii + jj) % 60).astype(f)
    A["vegmax"][...] = 80.0
    A["tsk"][...] = f(281.0) + f(0.25) * ((ii + jj) % 9).astype(f)
    snowh = np.where(ii % 4 == 0, f(0.3) + f(0.01) * jj.astype(f), f(0.0)).astype(f)
    A["snowh"][...] = snowh
    A["snow"][...] = f(250.0) * snowh
    for k in range(4):
        A["tslb"][:, k, :] = f(282.0) - f(0.5) * f(k)
        A["smois"][:, k, :] = f(0.20) + f(0.01) * ((ii + k) % 10).astype(f)
    for k in range(2):
        A["t3d"][:, k, :] = f(279.0) + f(0.2) * (jj % 5).astype(f)
        A["qv3d"][:, k, :] = 0.004
        A["u_phy"][:, k, :] = 3.0
        A["v_phy"][:, k, :] = -1.0
        A["p8w3d"][:, k, :] = 95000.0
        A["dz8w"][:, k, :] = 60.0
    A["coszin"][...] = 0.5
    A["swdown"][...] = 400.0
    A["glw"][...] = 300.0
    A["rainbl"][...] = np.where(ii % 3 == 0, 1.0, 0.0)
    A["dzs"] = S.DZS.copy()
    m = noahmp_b200.NoahMP(tables_usgs, ni, nj, device=0)
    I = {n: A[n] for n, k in _capi.INIT_SPEC if k in ("pf", "pi") and n in A}
    I["chstarxy"] = np.zeros((nj, ni), f)
    sc_i = dict(isurban=1, isice=24, iswater=16, fndsoilw=0, fndsnowh=1, nsoil=4, restart=0, allowed_to_read=1, iopt_run=1,
                dx=1000.0, dy=1000.0, wtddt=30.0, dt=3600.0, ids=1, ide=ni + 1, jds=1, jde=nj + 1, kds=1, kde=2, ims=1, ime=ni,
                jms=1, jme=nj, kms=1, kme=2, its=1, ite=ni, jts=1, jte=nj, kts=1, kte=2)
    m.init(I, sc_i)
    A["eahxy"][...] = (A["p8w3d"][:, 0, :] * f(0.004)) / (f(0.622) + f(0.004))
    A["tahxy"][...] = A["t3d"][:, 0, :]
    A["chxy"][...] = 0.1; A["cmxy"][...] = 0.1; A["albedo"][...] = 0.2; A["emiss"][...] = 0.95
    A["qsfc"][...] = f(0.004) / f(1.004)
    A["xlaixy"][...] = 1.0
    sc = dict(S.named_config("C1").opts)
    sc.update(yr=2017, julian=120.5, dt=3600.0, nsoil=4, dx=1000.0, xice_thres=0.5, isice=24, isurban=1, ids=1, ide=ni, jds=1,
              jde=nj, kds=1, kde=2, ims=1, ime=ni, jms=1, jme=nj, kms=1, kme=2, its=1, ite=ni, jts=1, jte=nj, kts=1, kte=2)
    for step in range(1, nsteps + 1):
        sc["itimestep"] = step
        assert m.noahmplsm(A, sc).code == 0
    m.close()
    for c in (0, 2 * ni + 4, 3 * ni + 2, ni * nj - 1):
        j, i = divmod(c, ni)
        for n in ("tsk", "hfx", "lh", "snow", "t2mbxy"):
            assert got[f"{n}[{c}]"] == "%08x" % int(A[n][j, i].view(np.uint32)), (n, c)
        assert int(got[f"isnowxy[{c}]"]) == int(A["isnowxy"][j, i])
    tot = (((A["tsk"] + A["hfx"]) + A["lh"]) + A["snow"]).astype(np.float64).sum()  # fp32 per cell, fp64 across cells
    assert float(got["checksum"]) == pytest.approx(tot, rel=1e-12)
    assert (A["isnowxy"] < 0).any() and np.isfinite(A["tsk"]).all()
