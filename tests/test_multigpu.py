"""Two GPUs, one process each (NCCL): tiles of one domain step independently and exchange only the opt_run=5
groundwater halo; the union of the tiles must equal the single-domain oracle bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, gni, gnj, nsteps, outdir):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import noahmp_b200
    from noahmp_b200 import halo, synthetic as S, tables
    td = tables.default_tables("USGS")
    cfg = S.named_config("C4"); cfg.ni, cfg.nj = gni, gnj
    cfg.opts["iopt_run"] = 5
    xs, xe, ys, ye = noahmp_b200.tile(gni, gnj, world, rank)
    ni, nj = xe - xs + 1, ye - ys + 1
    xp = S.backend()
    st = S.static_fields(xp, cfg, xs, xe, ys, ye)
    state = S.cold_start(cfg, st, S.forcing(xp, cfg, 1, st), td)
    wt, wsc = S.groundwater_fields(cfg, st, state)
    bounds = dict(ims=xs, ime=xe, its=xs, ite=xe, jms=ys, jme=ye, jts=ys, jte=ye, ide=gni, jde=gnj)
    wsc.update(bounds)
    m = noahmp_b200.NoahMP(td, ni, nj, device=rank, sync=noahmp_b200.SYNC_RESIDENT, math=noahmp_b200.MATH_PARITY)
    for step in range(1, nsteps + 1):
        arr, sc = S.args_from(cfg, st, S.forcing(xp, cfg, step, st), state, step)
        sc.update(bounds)
        s = m.noahmplsm(arr, sc)
        assert s.code == 0
        m.wtable_begin(wt, wsc)
        k, h = m.wtable_halo()
        halo.exchange_halo(torch.as_tensor(k, device=f"cuda:{rank}"), torch.as_tensor(h, device=f"cuda:{rank}"), rank, world)
        torch.cuda.synchronize()
        m.wtable_end(wt, wsc)
    m.sync_host(arr, sc)
    m.wtable_sync_host(wt, wsc)
    np.savez(os.path.join(outdir, f"gw{rank}.npz"), tile=np.array([xs, xe, ys, ye]), wtd=state["zwtxy"], smois=state["smois"],
             qslat=wt["qslat"], tsk=state["tsk"])
    m.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_two_gpu_groundwater_equals_single_domain(built, tables_usgs, tmp_path):
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from noahmp_b200 import _capi, synthetic as S
    from oracle import oracle as O
    from helpers import make_case
    gni, gnj, nsteps, world = 60, 44, 6, 2
    mp.spawn(_worker, args=(world, _free_port(), gni, gnj, nsteps, str(tmp_path)), nprocs=world, join=True)
    cfg = S.named_config("C4"); cfg.ni, cfg.nj = gni, gnj
    cfg.opts["iopt_run"] = 5
    ts = _capi.tables_from_dict(tables_usgs)
    _, st, state = make_case(cfg, tables_usgs)
    wt, wsc = S.groundwater_fields(cfg, st, state)
    xp = S.backend()
    O.set_math_mode(1)
    for step in range(1, nsteps + 1):
        arr, sc = S.args_from(cfg, st, S.forcing(xp, cfg, step, st), state, step)
        status, _ = O.noahmplsm(arr, sc, ts, nthreads=4)
        assert status.code == 0
        O.wtable(wt, wsc, ts)
    for r in range(world):
        z = np.load(os.path.join(tmp_path, f"gw{r}.npz"))
        xs, xe, ys, ye = z["tile"]
        assert np.array_equal(z["wtd"], state["zwtxy"][ys - 1:ye, xs - 1:xe])
        assert np.array_equal(z["qslat"], wt["qslat"][ys - 1:ye, xs - 1:xe])
        assert np.array_equal(z["smois"], state["smois"][ys - 1:ye, :, xs - 1:xe])
        assert np.array_equal(z["tsk"], state["tsk"][ys - 1:ye, xs - 1:xe])
    # the tile boundary carries flux: the halo mattered
    q = wt["qslat"]
    assert np.abs(q[:, gni // 2 - 1:gni // 2 + 1]).max() > 0


# ---- the exchanges inside the C library (its own NCCL communicator; torch.distributed is not involved) ----------------
def _lib_worker(rank, world, gni, gnj, nsteps, outdir):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import time
    torch.cuda.set_device(rank)
    import noahmp_b200
    from noahmp_b200 import synthetic as S, tables
    td = tables.default_tables("USGS")
    cfg = S.named_config("C4"); cfg.ni, cfg.nj = gni, gnj
    cfg.opts["iopt_run"] = 5
    xs, xe, ys, ye = noahmp_b200.tile(gni, gnj, world, rank)
    ni, nj = xe - xs + 1, ye - ys + 1
    xp = S.backend()
    st = S.static_fields(xp, cfg, xs, xe, ys, ye)
    state = S.cold_start(cfg, st, S.forcing(xp, cfg, 1, st), td)
    wt, wsc = S.groundwater_fields(cfg, st, state)
    bounds = dict(ims=xs, ime=xe, its=xs, ite=xe, jms=ys, jme=ye, jts=ys, jte=ye, ide=gni, jde=gnj)
    wsc.update(bounds)
    m = noahmp_b200.NoahMP(td, ni, nj, device=rank, sync=noahmp_b200.SYNC_RESIDENT, math=noahmp_b200.MATH_PARITY)
    # the host program only carries the 128-byte id from rank 0 to the others (a file here, MPI_Bcast in HRLDAS)
    idfile = os.path.join(outdir, "nccl_id.bin")
    if rank == 0:
        m.comm_unique_id().tofile(idfile + ".tmp")
        os.replace(idfile + ".tmp", idfile)
    while not os.path.exists(idfile):
        time.sleep(0.01)
    m.comm_init(np.fromfile(idfile, np.uint8), rank, world)
    assert m.comm_neighbours() == noahmp_b200.tile_neighbours(world, rank)
    m.budget_enable(True)
    for step in range(1, nsteps + 1):
        arr, sc = S.args_from(cfg, st, S.forcing(xp, cfg, step, st), state, step)
        sc.update(bounds)
        assert m.noahmplsm(arr, sc).code == 0
        m.wtable(wt, wsc)  # begin + NCCL halo + end, all inside the library
    local = m.budget_read()
    glob = m.budget_read(global_sum=True)
    m.sync_host(arr, sc)
    m.wtable_sync_host(wt, wsc)
    np.savez(os.path.join(outdir, f"lib{rank}.npz"), tile=np.array([xs, xe, ys, ye]), wtd=state["zwtxy"], smois=state["smois"],
             qslat=wt["qslat"], tsk=state["tsk"], local=np.array([local[k] for k in m.BUDGET_NAMES]),
             glob=np.array([glob[k] for k in m.BUDGET_NAMES]))
    m.close()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_library_nccl_halo_and_global_budget(built, tables_usgs, tmp_path, world):
    """2 GPUs (a 2x1 process grid: columns only), 4 GPUs (2x2: rows and corners) and 8 GPUs (4x2, the grid the CONUS
    bench runs on; interior tiles have neighbours on three sides): noahmp_b200_wtable exchanges the
    KCELL / HEAD halo itself over NCCL and the union of the tiles equals the single-domain oracle bit for bit; the
    all-reduced budget equals the sum of the tiles' own sums."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs (gpurun --gpus {world})")
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from noahmp_b200 import _capi, synthetic as S
    from oracle import oracle as O
    from helpers import make_case
    gni, gnj, nsteps = 61, 45, 6
    mp.spawn(_lib_worker, args=(world, gni, gnj, nsteps, str(tmp_path)), nprocs=world, join=True)
    cfg = S.named_config("C4"); cfg.ni, cfg.nj = gni, gnj
    cfg.opts["iopt_run"] = 5
    ts = _capi.tables_from_dict(tables_usgs)
    _, st, state = make_case(cfg, tables_usgs)
    wt, wsc = S.groundwater_fields(cfg, st, state)
    xp = S.backend()
    O.set_math_mode(1)
    for step in range(1, nsteps + 1):
        arr, sc = S.args_from(cfg, st, S.forcing(xp, cfg, step, st), state, step)
        status, _ = O.noahmplsm(arr, sc, ts, nthreads=4)
        assert status.code == 0
        O.wtable(wt, wsc, ts)
    tot = np.zeros(8)
    zs = [np.load(os.path.join(tmp_path, f"lib{r}.npz")) for r in range(world)]
    for z in zs:
        xs, xe, ys, ye = z["tile"]
        assert np.array_equal(z["wtd"], state["zwtxy"][ys - 1:ye, xs - 1:xe])
        assert np.array_equal(z["qslat"], wt["qslat"][ys - 1:ye, xs - 1:xe])
        assert np.array_equal(z["smois"], state["smois"][ys - 1:ye, :, xs - 1:xe])
        assert np.array_equal(z["tsk"], state["tsk"][ys - 1:ye, xs - 1:xe])
        tot += z["local"]
    tot[7] = nsteps
    for z in zs:
        assert np.allclose(z["glob"], tot, rtol=1e-12, atol=1e-9), (z["glob"], tot)
    q = wt["qslat"]
    assert np.abs(q[:, gni // 2 - 1:gni // 2 + 1]).max() > 0  # flux crosses the tile boundary: the halo mattered
    if world >= 4:
        assert np.abs(q[gnj // 2 - 1:gnj // 2 + 1, gni // 2 - 1:gni // 2 + 1]).max() > 0  # and the corner cells
