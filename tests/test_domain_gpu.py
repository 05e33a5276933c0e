"""Row f4 — one host process, several GPUs: the per-tile contexts slice their tile out of whole-domain host arrays
(memory bounds = the domain, tile bounds = the GPU's tile), replacing the IO-rank scatter / gather of the reference
(decompose_data_*, write_io_*, mpp/module_mpp_land.F90:645-857).  On a one-GPU box the tiles share the GPU; with more
GPUs they spread over them.  The result must equal the single-tile run of the same domain bit for bit."""
import numpy as np
import pytest
import torch

from noahmp_b200 import _capi, synthetic as S

from helpers import clone_state, diff_report, make_case, run_oracle

pytestmark = pytest.mark.gpu


def _cfg(name, ni, nj, **opts):
    cfg = S.named_config(name)
    cfg.ni, cfg.nj = ni, nj
    cfg.opts.update(opts)
    return cfg


@pytest.mark.parametrize("ntiles,sync", [(2, 0), (4, 0), (4, 1), (6, 1)])
def test_domain_of_tiles_equals_single_tile(built, tables_usgs, ntiles, sync):
    import noahmp_b200
    cfg = _cfg("C4", 101, 67)  # not divisible: tiles of unequal width and height
    ts = _capi.tables_from_dict(tables_usgs)
    _, st, state0 = make_case(cfg, tables_usgs)
    st["xice"][20:23, 30:60] = 1.0
    ref, dom = clone_state(state0), clone_state(state0)
    ngpu = torch.cuda.device_count()
    devices = [r % ngpu for r in range(ntiles)]
    fetch = ["tsk", "tslb", "isnowxy"] if sync else None
    D = noahmp_b200.NoahMPDomain(tables_usgs, cfg.ni, cfg.nj, devices, sync=sync, math=noahmp_b200.MATH_PARITY,
                                 fetch=fetch)
    assert D.ntiles == ntiles
    seen = np.zeros((cfg.nj, cfg.ni), int)
    for r in range(ntiles):
        xs, xe, ys, ye = D.tile_bounds(r)
        assert (xs, xe, ys, ye) == noahmp_b200.tile(cfg.ni, cfg.nj, ntiles, r)
        seen[ys - 1:ye, xs - 1:xe] += 1
    assert (seen == 1).all()
    assert run_oracle(cfg, ts, st, ref, 5, math_mode=1) is None
    xp = S.backend()
    for step in range(1, 6):
        arr, sc = S.args_from(cfg, st, S.forcing(xp, cfg, step, st), dom, step)
        s = D.noahmplsm(arr, sc)
        assert (s.code, s.count) == (0, 0)
        if sync:
            for n in fetch:
                pass  # compared below for the last step
    if sync:
        for n in fetch:
            assert np.array_equal(ref[n], dom[n]), n  # the fetch list is current after every call
        assert diff_report(ref, dom, ["hfx", "smois"])  # the rest waits for sync_host
        D.sync_host(arr, sc)
    rep = diff_report(ref, dom)
    assert not rep, rep
    D.close()


def test_domain_status_is_the_first_failing_column_in_loop_order(built, tables_usgs):
    import noahmp_b200
    cfg = _cfg("C1", 40, 30)
    ts = _capi.tables_from_dict(tables_usgs)
    _, st, state0 = make_case(cfg, tables_usgs)
    st["isltyp"][22, 31] = 0   # Fortran (32, 23): in the upper right tile of a 2x2 grid
    st["isltyp"][9, 35] = 25   # Fortran (36, 10): lower right tile -> comes first in the j-outer loop order
    ref, dom = clone_state(state0), clone_state(state0)
    e = run_oracle(cfg, ts, st, ref, 1, math_mode=1)
    D = noahmp_b200.NoahMPDomain(tables_usgs, cfg.ni, cfg.nj, [0, 0, 0, 0], math=noahmp_b200.MATH_PARITY)
    xp = S.backend()
    arr, sc = S.args_from(cfg, st, S.forcing(xp, cfg, 1, st), dom, 1)
    s = D.noahmplsm(arr, sc)
    assert (s.code, s.i, s.j, s.count) == (e[1], e[2], e[3], e[4]) == (7, 36, 10, 2)
    assert not diff_report(ref, dom)
    D.close()


def test_memory_bounds_larger_than_the_tile(built, tables_usgs):
    """The per-tile call itself accepts host arrays that extend beyond its tile (ims < its ...): the cells outside the
    tile are neither read nor written."""
    import noahmp_b200
    cfg = _cfg("C4", 90, 60)
    ts = _capi.tables_from_dict(tables_usgs)
    _, st, state0 = make_case(cfg, tables_usgs)
    ref, big = clone_state(state0), clone_state(state0)
    assert run_oracle(cfg, ts, st, ref, 3, math_mode=1) is None
    its, ite, jts, jte = 11, 70, 6, 50
    m = noahmp_b200.NoahMP(tables_usgs, ite - its + 1, jte - jts + 1, device=0, math=noahmp_b200.MATH_PARITY)
    xp = S.backend()
    for step in (1, 2, 3):
        arr, sc = S.args_from(cfg, st, S.forcing(xp, cfg, step, st), big, step)
        sc.update(its=its, ite=ite, jts=jts, jte=jte)
        assert m.noahmplsm(arr, sc).code == 0
    m.close()
    inside = np.zeros((cfg.nj, cfg.ni), bool)
    inside[jts - 1:jte, its - 1:ite] = True
    for n in ("tsk", "hfx", "isnowxy", "t2mvxy"):
        assert np.array_equal(big[n][inside], ref[n][inside]), n
        assert np.array_equal(big[n][~inside], state0[n][~inside]), n
    m3 = np.broadcast_to(inside[:, None, :], big["tslb"].shape)
    assert np.array_equal(big["tslb"][m3], ref["tslb"][m3]) and np.array_equal(big["tslb"][~m3], state0["tslb"][~m3])
