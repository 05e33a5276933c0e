"""opt_run = 5 (Miguez-Macho & Fan groundwater): oracle known answers for LATERALFLOW / UPDATEWTD and the halo
exchange logic on CPU (gloo, world_size 2 and 4).  GPU parity of the same path is in test_parity_gpu.py."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from noahmp_b200 import _capi, synthetic as S

from helpers import clone_state, make_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _case(tables, ni=40, nj=30, name="C4"):
    cfg = S.named_config(name); cfg.ni, cfg.nj = ni, nj
    cfg.opts["iopt_run"] = 5
    _, st, state = make_case(cfg, tables)
    wt, wsc = S.groundwater_fields(cfg, st, state)
    return cfg, st, state, wt, wsc


def test_flat_head_gives_no_lateral_flow(built, tables_usgs):
    from oracle import oracle as O
    cfg, st, state, wt, wsc = _case(tables_usgs)
    ts = _capi.tables_from_dict(tables_usgs)
    wt["topo"][...] = 100.0
    wt["wtd"][...] = -5.0
    wt["rivercond"][...] = 0.0
    O.wtable(wt, wsc, ts)
    assert np.abs(wt["qslat"]).max() == 0.0 and wt["qrf"].max() == 0.0


def test_mound_drains_to_its_neighbours_antisymmetrically(built, tables_usgs):
    """A single raised water table on flat, homogeneous land: the centre loses water, its 8 neighbours gain, and
    the flux the centre sends to a neighbour equals what that neighbour receives (uniform AREA)."""
    from oracle import oracle as O
    cfg, st, state, wt, wsc = _case(tables_usgs, name="C1")
    ts = _capi.tables_from_dict(tables_usgs)
    st["isltyp"][...] = 3; wt["isltyp"][...] = 3
    wt["topo"][...] = 100.0; wt["wtd"][...] = -6.0; wt["fdepth"][...] = 100.0; wt["rivercond"][...] = 0.0
    wt["wtd"][15, 20] = -2.0
    O.wtable(wt, wsc, ts)
    q = wt["qslat"]
    assert q[15, 20] < 0
    ring = [q[15 + dj, 20 + di] for dj in (-1, 0, 1) for di in (-1, 0, 1) if (dj, di) != (0, 0)]
    assert min(ring) > 0
    assert abs(q[15, 20] + sum(ring)) <= 1e-4 * abs(q[15, 20])
    assert abs(q[15, 19] - q[15, 21]) <= 1e-6 * abs(q[15, 19]) and q[15, 19] > q[14, 19]  # axial > diagonal
    far = q.copy(); far[14:17, 19:22] = 0
    assert np.abs(far).max() == 0.0


def test_reference_index_quirk_of_the_per_step_call(built, tables_usgs):
    """With IDE=xend, JDE=yend (module_hrldas_noahmp_driver.F90:432) fluxes exist only on 2..NI-2 x 2..NJ-2
    (SURVEY.md Appendix A #20)."""
    from oracle import oracle as O
    cfg, st, state, wt, wsc = _case(tables_usgs, name="C1")
    ts = _capi.tables_from_dict(tables_usgs)
    wt["rivercond"][...] = 0.0
    O.wtable(wt, wsc, ts)
    q = wt["qslat"]
    assert np.abs(q[1:-2, 1:-2]).max() > 0
    assert np.abs(q[0, :]).max() == 0 and np.abs(q[-2:, :]).max() == 0
    assert np.abs(q[:, 0]).max() == 0 and np.abs(q[:, -2:]).max() == 0


def test_updatewtd_conserves_column_water(built, tables_usgs):
    """Rising water table (TOTWATER > 0, :343-470): soil-water change + deep-layer change + spring = TOTWATER.
    (The falling branch is not mass-conserving in the reference itself: `smc = smc + maxwatdw/dzs` at :520 adds
    the water a layer yields instead of removing it; the oracle restates it as written.)"""
    from oracle import oracle as O
    cfg, st, state, wt, wsc = _case(tables_usgs, name="C1")
    ts = _capi.tables_from_dict(tables_usgs)
    wt["rivercond"][...] = 0.0
    wt["wtd"][...] = -1.2                      # inside layer 4 (-1.0 .. -2.0)
    wt["wtd"][10:20, 10:30] = -0.9             # a plateau: water leaves it, arrives next to it
    dz = S.DZS
    before = (state["smois"] * dz[None, :, None]).sum(1)
    smcwtd0 = wt["smcwtd"].copy()
    O.wtable(wt, wsc, ts)
    after = (state["smois"] * dz[None, :, None]).sum(1)
    qlat = wt["qslat"] * 1e-3
    moved = (after - before) + wt["qspring"] + (wt["smcwtd"] - smcwtd0) * dz[-1]
    sel = qlat > 0
    assert sel.sum() > 30
    assert np.abs(moved - qlat)[sel].max() < 2e-6


def test_coupled_run_stays_sane(built, tables_usgs):
    """24 hourly steps of noahmplsm(opt_run=5) + WTABLE every step on a mixed land/water/glacier tile."""
    from oracle import oracle as O
    cfg, st, state, wt, wsc = _case(tables_usgs, 48, 40)
    ts = _capi.tables_from_dict(tables_usgs)
    xp = S.backend()
    O.set_math_mode(0)
    for step in range(1, 25):
        arr, sc = S.args_from(cfg, st, S.forcing(xp, cfg, step, st), state, step)
        status, _ = O.noahmplsm(arr, sc, ts, nthreads=4)
        assert status.code == 0, (step, status.code, status.i, status.j, status.value)
        O.wtable(wt, wsc, ts)
    land = (st["xland"] < 1.5) & (st["ivgtyp"] != S.ISICE)
    assert np.isfinite(state["zwtxy"]).all() and np.isfinite(state["smois"]).all()
    assert (state["smois"][np.broadcast_to(land[:, None, :], state["smois"].shape)] > 0).all()
    assert (state["deeprechxy"] == 0).all() and np.abs(state["rechxy"][land]).max() > 0
    assert wt["qrfs"][land].min() >= 0


# ---- halo exchange logic (gloo) ---------------------------------------------------------------------------------
def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _halo_worker(rank, world, port, gni, gnj, outdir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import noahmp_b200
    from noahmp_b200 import halo
    xs, xe, ys, ye = noahmp_b200.tile(gni, gnj, world, rank)
    ni, nj = xe - xs + 1, ye - ys + 1
    jj, ii = np.meshgrid(np.arange(ys, ye + 1), np.arange(xs, xe + 1), indexing="ij")
    k = torch.zeros((nj + 2, ni + 2)); h = torch.zeros((nj + 2, ni + 2))
    k[1:-1, 1:-1] = torch.from_numpy((ii + 1000 * jj).astype(np.float32))      # value = global cell id
    h[1:-1, 1:-1] = torch.from_numpy((-(ii + 1000 * jj)).astype(np.float32))
    halo.exchange_halo(k, h, rank, world)
    np.savez(os.path.join(outdir, f"halo{rank}.npz"), k=k.numpy(), h=h.numpy(), tile=np.array([xs, xe, ys, ye]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 6])
def test_halo_exchange_delivers_edges_and_corners(built, tmp_path, world):
    gni, gnj = 23, 17
    mp.spawn(_halo_worker, args=(world, _free_port(), gni, gnj, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        z = np.load(os.path.join(tmp_path, f"halo{r}.npz"))
        xs, xe, ys, ye = z["tile"]
        for jl in range(z["k"].shape[0]):
            for il in range(z["k"].shape[1]):
                I, J = xs - 1 + il, ys - 1 + jl
                inside = 1 <= I <= gni and 1 <= J <= gnj
                want = float(I + 1000 * J) if inside else 0.0
                assert z["k"][jl, il] == want, (r, il, jl)
                assert z["h"][jl, il] == -want
