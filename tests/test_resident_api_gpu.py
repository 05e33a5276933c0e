"""RESIDENT-mode contract of the drop-in call (state in HBM across calls): what travels per call, what may be declared
unchanged, what the host may rewrite between calls, how the row-chunk pipeline and the re-binning interact, the status
latch of device-side loops, the page-lock budget, and the global budget sums.  All through the C-ABI."""
import numpy as np
import pytest

from noahmp_b200 import _capi, synthetic as S

from helpers import clone_state, diff_report, make_case, run_oracle

pytestmark = pytest.mark.gpu


def _cfg(name, ni, nj, **opts):
    cfg = S.named_config(name)
    cfg.ni, cfg.nj = ni, nj
    cfg.opts.update(opts)
    return cfg


def _model(tables, cfg, sync, math=None):
    import noahmp_b200
    return noahmp_b200.NoahMP(tables, cfg.ni, cfg.nj, device=0, sync=sync,
                              math=noahmp_b200.MATH_PARITY if math is None else math)


def test_forcing_hints_do_not_change_results(built, tables_usgs):
    """DZ8W constant, VEGFRA unchanged, P8W3D levels equal: with the hints the three planes are sent once and the
    state ends bit-identical to the run that sends all twelve planes every call; withdrawing the VEGFRA hint for the
    call after a change makes the new values count."""
    import noahmp_b200
    cfg = _cfg("C4", 120, 90)
    _, st, state0 = make_case(cfg, tables_usgs)
    a, b = clone_state(state0), clone_state(state0)
    m1 = _model(tables_usgs, cfg, noahmp_b200.SYNC_RESIDENT)
    m2 = _model(tables_usgs, cfg, noahmp_b200.SYNC_RESIDENT)
    allh = noahmp_b200.HINT_DZ8W_CONSTANT | noahmp_b200.HINT_VEGFRA_UNCHANGED | noahmp_b200.HINT_P8W_LEVELS_EQUAL
    m2.set_forcing_hints(allh)
    xp = S.backend()
    vegfra2 = (st["vegfra"] * np.float32(0.5)).astype(np.float32)
    for step in range(1, 9):
        frc = S.forcing(xp, cfg, step, st)
        arr_a, sc = S.args_from(cfg, st, frc, a, step)
        arr_b, _ = S.args_from(cfg, st, frc, b, step)
        if step >= 5:  # the driver read a new VEGFRA before step 5
            arr_a["vegfra"] = vegfra2
            arr_b["vegfra"] = vegfra2
        if step == 5:
            m2.set_forcing_hints(allh & ~noahmp_b200.HINT_VEGFRA_UNCHANGED)
        if step == 6:
            m2.set_forcing_hints(allh)
        if step >= 2:  # a wrong DZ8W / level 2 on the hinted side must not be read any more
            arr_b["dz8w"] = np.full_like(arr_b["dz8w"], 1.0e3)
            arr_b["p8w3d"] = arr_b["p8w3d"].copy()
            arr_b["p8w3d"][:, 1, :] = 5.0e4
        s1, s2 = m1.noahmplsm(arr_a, sc), m2.noahmplsm(arr_b, sc)
        assert (s1.code, s2.code) == (0, 0)
    arr_a, sc = S.args_from(cfg, st, S.forcing(xp, cfg, 8, st), a, 8)
    arr_b, _ = S.args_from(cfg, st, S.forcing(xp, cfg, 8, st), b, 8)
    m1.sync_host(arr_a, sc); m2.sync_host(arr_b, sc)
    rep = diff_report(a, b)
    assert not rep, rep
    m1.close(); m2.close()


def test_push_list_follows_host_lai(built, tables_usgs):
    """land_driver_exe overwrites LAI (XLAIXY) from the forcing file before every call: with dveg=2 the resident state
    must take the host's array again each step (set_push), as the strict drop-in mode does."""
    import noahmp_b200
    cfg = _cfg("C3", 96, 64)
    _, st, state0 = make_case(cfg, tables_usgs)
    a, b, c = clone_state(state0), clone_state(state0), clone_state(state0)
    m1 = _model(tables_usgs, cfg, noahmp_b200.SYNC_FULL)
    m2 = _model(tables_usgs, cfg, noahmp_b200.SYNC_RESIDENT)
    m3 = _model(tables_usgs, cfg, noahmp_b200.SYNC_RESIDENT)
    m2.set_push(["xlaixy"]); m2.set_fetch(["xlaixy"]); m2.set_chunks(3)
    xp = S.backend()
    for step in range(1, 7):
        frc = S.forcing(xp, cfg, step, st)
        lai = (1.0 + 0.5 * np.sin(0.7 * step) + 0.002 * (st["_g"] % 97)).astype(np.float32)  # "read from the forcing file"
        for s in (a, b, c):
            s["xlaixy"][...] = lai
        arr_a, sc = S.args_from(cfg, st, frc, a, step)
        arr_b, _ = S.args_from(cfg, st, frc, b, step)
        arr_c, _ = S.args_from(cfg, st, frc, c, step)
        assert m1.noahmplsm(arr_a, sc).code == 0 and m2.noahmplsm(arr_b, sc).code == 0 and m3.noahmplsm(arr_c, sc).code == 0
        assert np.array_equal(a["xlaixy"], b["xlaixy"]), step
    m2.sync_host(arr_b, sc); m3.sync_host(arr_c, sc)
    assert not diff_report(a, b)
    assert diff_report(a, c)  # without the push list the resident LAI evolves on its own
    with pytest.raises(noahmp_b200.NoahmpError):
        m2.set_push(["t2mvxy"])  # an OUT array cannot be pushed
    for m in (m1, m2, m3):
        m.close()


def test_rebinning_is_skipped_while_the_order_holds(built, tables_usgs, monkeypatch):
    """A due re-binning first counts the places where the bin keys decrease along the compact order and permutes only
    if more than a threshold of the columns left their bin: on a snow-free tile nothing ever does, so after the first
    binning no further permutation happens -- and with the threshold at 0 one happens at every interval.  Either way
    the results equal the FULL-mode results bit for bit."""
    import noahmp_b200
    cfg = _cfg("C2", 96, 80)
    _, st, state0 = make_case(cfg, tables_usgs)
    xp = S.backend()
    counts = {}
    ref = None
    for thr in ("default", "0"):
        if thr == "0":
            monkeypatch.setenv("NOAHMP_B200_REBIN_MIN_CHANGED", "0")
        s = clone_state(state0)
        m = _model(tables_usgs, cfg, noahmp_b200.SYNC_RESIDENT)
        m.set_rebin(3)
        m.set_fetch(["tsk", "hfx"])
        for step in range(1, 17):
            arr, sc = S.args_from(cfg, st, S.forcing(xp, cfg, step, st), s, step)
            assert m.noahmplsm(arr, sc).code == 0
        counts[thr] = m.rebins
        m.sync_host(arr, sc)
        if ref is None:
            ref = s
        else:
            assert not diff_report(ref, s)
        m.close()
    assert counts["default"] == 1 and counts["0"] >= 4, counts


def test_chunking_may_change_after_rebinning(built, tables_usgs, monkeypatch):
    """After the first re-binning the columns stay inside the row chunk they were binned in; calls that ask for another
    chunk count (a different entry point, set_chunks) keep working and give the same bits."""
    import noahmp_b200
    monkeypatch.setenv("NOAHMP_B200_REBIN_MIN_CHANGED", "0")   # permute at every interval, whether the order broke or not
    cfg = _cfg("C4", 150, 121)
    _, st, state0 = make_case(cfg, tables_usgs)
    a, b = clone_state(state0), clone_state(state0)
    m1 = _model(tables_usgs, cfg, noahmp_b200.SYNC_FULL)
    m2 = _model(tables_usgs, cfg, noahmp_b200.SYNC_RESIDENT)
    m2.set_rebin(3)
    m2.set_fetch(["tsk"])
    xp = S.backend()
    plan = {1: 4, 2: 4, 3: 4, 4: 4, 5: 1, 6: 7, 7: 2, 8: 4, 9: 1, 10: 5}
    import torch
    for step in range(1, 11):
        frc = S.forcing(xp, cfg, step, st)
        arr_a, sc = S.args_from(cfg, st, frc, a, step)
        arr_b, _ = S.args_from(cfg, st, frc, b, step)
        assert m1.noahmplsm(arr_a, sc).code == 0
        m2.set_chunks(plan[step])
        if step in (6, 9):  # the device-side entry point in between
            dev = [torch.as_tensor(x, device="cuda") for x in m2.device_forcing()]
            order = ["coszin", "t3d", "qv3d", "u_phy", "v_phy", "swdown", "glw", "p8w3d", "p8w3d", "rainbl", "vegfra", "dz8w"]
            for k, n in enumerate(order):
                h = arr_b[n]
                dev[k].copy_(torch.from_numpy(np.ascontiguousarray(h[:, 1 if k == 8 else 0, :] if h.ndim == 3 else h)))
            torch.cuda.synchronize()
            m2.bind_forcing(None)
            m2.step_device(step, sc["yr"], sc["julian"], sc["dt"])
            m2.fetch(arr_b, sc, "tsk")
        else:
            assert m2.noahmplsm(arr_b, sc).code == 0
        assert np.array_equal(a["tsk"], b["tsk"]), step
    assert m2.rebins >= 2
    m2.sync_host(arr_b, sc)
    assert not diff_report(a, b)
    m1.close(); m2.close()


def test_status_latch_keeps_first_failure_of_a_device_loop(built, tables_usgs):
    """A loop of step_device calls without host round trips: get_status afterwards reports the failures of ALL steps
    since the previous read (count summed, first failing column kept), then starts a new period."""
    import noahmp_b200
    cfg = _cfg("C1", 16, 12)
    _, st, state0 = make_case(cfg, tables_usgs)
    st["isltyp"][5, 3] = 0  # REDPRM range violation in one column, every step
    m = _model(tables_usgs, cfg, noahmp_b200.SYNC_RESIDENT)
    xp = S.backend()
    arr, sc = S.args_from(cfg, st, S.forcing(xp, cfg, 1, st), state0, 1)
    m.upload(arr, sc)
    for step in (1, 2, 3):
        m.step_device(step, sc["yr"], sc["julian"], sc["dt"])
    s = m.status()
    assert (s.code, s.i, s.j, s.count) == (7, 4, 6, 3)
    s = m.status()
    assert (s.code, s.count) == (0, 0)
    m.step_device(4, sc["yr"], sc["julian"], sc["dt"])
    assert m.status().count == 1
    m.close()


def test_full_sync_returns_the_callers_out_arrays_at_water_cells(built, tables_usgs):
    """Strict drop-in mode: OUT arrays at open-water cells come back as the caller holds them NOW (the reference
    never touches them), also when the caller changed them after the first call."""
    import noahmp_b200
    cfg = _cfg("C4", 64, 48)
    _, st, state = make_case(cfg, tables_usgs)
    m = _model(tables_usgs, cfg, noahmp_b200.SYNC_FULL)
    xp = S.backend()
    water = st["xland"] >= 1.5
    assert water.any()
    for step in (1, 2, 3):
        state["t2mvxy"][water] = 100.0 + step
        state["chb2xy"][water] = -7.0 * step
        arr, sc = S.args_from(cfg, st, S.forcing(xp, cfg, step, st), state, step)
        assert m.noahmplsm(arr, sc).code == 0
        assert (state["t2mvxy"][water] == np.float32(100.0 + step)).all()
        assert (state["chb2xy"][water] == np.float32(-7.0 * step)).all()
    m.close()


def test_page_lock_budget_is_bounded(built, tables_usgs, monkeypatch):
    """A driver loop that passes fresh arrays to every call does not grow the page-locked set without bound."""
    import noahmp_b200
    from noahmp_b200 import driver
    monkeypatch.setenv("NOAHMP_B200_PIN_BUDGET_GB", "0.25")
    monkeypatch.setattr(driver, "HOLD_BUDGET_BYTES", 256 << 20)
    cfg = _cfg("C2", 1200, 1000)  # 4.8 MB planes: above the 4 MiB page-lock threshold
    _, st, state0 = make_case(cfg, tables_usgs)
    a = clone_state(state0)
    m = _model(tables_usgs, cfg, noahmp_b200.SYNC_RESIDENT, math=noahmp_b200.MATH_FAST)
    xp = S.backend()
    base = None
    for step in range(1, 13):
        frc = S.forcing(xp, cfg, step, st)
        arr, sc = S.args_from(cfg, st, frc, a, step)  # fresh forcing arrays every step
        assert m.noahmplsm(arr, sc).code == 0
        if step == 2:
            base = m._held_bytes  # the state arrays (passed every call) + one call's forcing
        if step > 2:
            assert m._held_bytes <= base + 4 * cfg.ni * cfg.nj, (step, m._held_bytes, base)
    m.close()


def test_budget_sums_match_numpy(built, tables_usgs):
    """Row e3 on one tile: the eight fp64 budget sums against numpy over the oracle's arrays; the water residual of the
    interval is the sum of the per-column ERRWAT, far below the 0.1 mm per column the model tolerates."""
    import noahmp_b200
    cfg = _cfg("C4", 96, 72)
    ts = _capi.tables_from_dict(tables_usgs)
    _, st, state0 = make_case(cfg, tables_usgs)
    s_cpu, s_gpu = clone_state(state0), clone_state(state0)
    m = _model(tables_usgs, cfg, noahmp_b200.SYNC_RESIDENT)
    m.budget_enable(True)
    xp = S.backend()
    nonwater = st["xland"] < 1.5
    glac = nonwater & (st["ivgtyp"] == S.ISICE)
    land = nonwater & ~glac
    dz = S.DZS.astype(np.float64)

    def storage(s):
        soil = (s["smois"].astype(np.float64) * dz[None, :, None]).sum(axis=1) * 1000.0
        sto = s["snow"].astype(np.float64).copy()
        sto[land] += (s["canwat"].astype(np.float64) + s["waxy"] + soil)[land]
        return sto[nonwater].sum()

    acc = dict(precip=0.0, et=0.0, runoff=0.0, erreng=0.0)
    sto0 = None
    for step in range(1, 7):
        frc = S.forcing(xp, cfg, step, st)
        a1, sc = S.args_from(cfg, st, frc, s_cpu, step)
        a2, _ = S.args_from(cfg, st, frc, s_gpu, step)
        from oracle import oracle as O
        O.set_math_mode(1)
        e1, _ = O.noahmplsm(a1, sc, ts, nthreads=4)
        assert e1.code == 0 and m.noahmplsm(a2, sc).code == 0
        f64 = lambda n: s_cpu[n].astype(np.float64)
        acc["precip"] += frc["rainbl"].astype(np.float64)[nonwater].sum()
        acc["et"] += ((f64("ecanxy") + f64("edirxy") + f64("etranxy")) * cfg.dt)[nonwater].sum()
        acc["runoff"] += ((f64("runsfxy") + f64("runsbxy")) * cfg.dt)[nonwater].sum()
        sav = np.where(land, f64("savxy"), 0.0)
        acc["erreng"] += (sav + f64("sagxy") - (f64("firaxy") + f64("hfx") + f64("lh") + f64("grdflx")))[nonwater].sum()
        if step == 1:
            b1 = m.budget_read()
            sto0 = storage(state0)
    b = m.budget_read(global_sum=True)  # no communicator: the tile's own sums
    rel = lambda x, y: abs(x - y) / max(abs(y), 1.0)
    assert rel(b["storage_mm"], storage(s_cpu)) < 1e-9
    assert rel(b["precip_mm"], acc["precip"]) < 1e-9 and rel(b["et_mm"], acc["et"]) < 1e-9
    assert rel(b["runoff_mm"], acc["runoff"]) < 1e-9
    assert abs(b["erreng_wm2"] - acc["erreng"]) < 1e-6 * nonwater.sum()
    assert rel(b["swe_mm"], s_cpu["snow"].astype(np.float64)[nonwater].sum()) < 1e-9
    assert (b["columns"], b["steps"]) == (float(nonwater.sum()), 6.0) and b1["steps"] == 1.0
    # global water balance of the interval (land columns carry the full balance; the glacier soil is ice by definition)
    resid = (b["storage_mm"] - sto0) - (b["precip_mm"] - b["et_mm"] - b["runoff_mm"])
    assert abs(resid) < 0.1 * nonwater.sum() * 6
    assert abs(b["erreng_wm2"]) < 0.01 * nonwater.sum() * 6
    m.budget_read(reset=True)
    assert m.budget_read()["steps"] == 0.0
    m.close()
