"""The N>1 path on CPU: two `gloo` ranks each own an mpp_land tile of one domain and step it independently (the
column physics has no exchange step), then reduce what bench.py reduces (column counts by SUM, time by MAX).
Since no GPU exists here the per-tile stepping uses the CPU oracle; what is under test is the host-side logic
shared with the GPU path: tile maps, tile-local synthetic inputs keyed by the GLOBAL column index, memory-bound
bookkeeping of the `noahmplsm` arguments and the cross-rank reductions."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, ni, nj, nsteps, outdir):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import noahmp_b200
    from noahmp_b200 import _capi, synthetic as S, tables
    from oracle import oracle as O
    td = tables.default_tables("USGS"); ts = _capi.tables_from_dict(td)
    cfg = S.named_config("C4"); cfg.ni, cfg.nj = ni, nj
    xs, xe, ys, ye = noahmp_b200.tile(ni, nj, world, rank)
    xp = S.backend()
    st = S.static_fields(xp, cfg, xs, xe, ys, ye)
    state = S.cold_start(cfg, st, S.forcing(xp, cfg, 1, st), td)
    O.set_math_mode(1)
    for step in range(1, nsteps + 1):
        arr, sc = S.args_from(cfg, st, S.forcing(xp, cfg, step, st), state, step)
        sc.update(ims=xs, ime=xe, its=xs, ite=xe, jms=ys, jme=ye, jts=ys, jte=ye, ide=ni, jde=nj)
        status, _ = O.noahmplsm(arr, sc, ts, nthreads=1)
        assert status.code == 0
    ncol = int((st["xland"] < 1.5).sum())
    v = torch.tensor([float(ncol), float(rank + 1)], dtype=torch.float64)
    tot = v.clone(); dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    mx = v.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    np.savez(os.path.join(outdir, f"rank{rank}.npz"), tile=np.array([xs, xe, ys, ye]), tsk=state["tsk"],
             tslb=state["tslb"], isnow=state["isnowxy"], total_cols=tot[0].item(), max_t=mx[1].item())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_tiled_run_equals_single_domain(built, tables_usgs, tmp_path, world):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from noahmp_b200 import _capi, synthetic as S
    from helpers import make_case, run_oracle
    ni, nj, nsteps = 50, 37, 3
    mp.spawn(_worker, args=(world, _free_port(), ni, nj, nsteps, str(tmp_path)), nprocs=world, join=True)
    cfg = S.named_config("C4"); cfg.ni, cfg.nj = ni, nj
    ts = _capi.tables_from_dict(tables_usgs)
    _, st, state = make_case(cfg, tables_usgs)
    assert run_oracle(cfg, ts, st, state, nsteps, math_mode=1, nthreads=2) is None
    ncol = int((st["xland"] < 1.5).sum())
    seen = np.zeros((nj, ni), bool)
    for r in range(world):
        z = np.load(os.path.join(tmp_path, f"rank{r}.npz"))
        xs, xe, ys, ye = z["tile"]
        assert z["total_cols"] == ncol and z["max_t"] == world
        assert np.array_equal(z["tsk"], state["tsk"][ys - 1:ye, xs - 1:xe])
        assert np.array_equal(z["tslb"], state["tslb"][ys - 1:ye, :, xs - 1:xe])
        assert np.array_equal(z["isnow"], state["isnowxy"][ys - 1:ye, xs - 1:xe])
        seen[ys - 1:ye, xs - 1:xe] = True
    assert seen.all()


def _budget_worker(rank, world, port, outdir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from noahmp_b200 import halo
    names = ("storage_mm", "precip_mm", "et_mm", "runoff_mm", "erreng_wm2", "swe_mm", "columns", "steps")
    local = {n: float((rank + 1) * (k + 1)) for k, n in enumerate(names)}
    local["steps"] = 24.0
    out = halo.allreduce_budget(local)
    np.save(os.path.join(outdir, f"b{rank}.npy"), np.array([out[n] for n in names]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_budget_allreduce_over_gloo(tmp_path, world):
    """Row e3, host-side form: eight fp64 budget sums per tile, one all-reduce (SUM); every rank ends with the same
    global sums and the common step count.  (On the GPU box the library does the same with one ncclAllReduce:
    tests/test_multigpu.py.)"""
    mp.spawn(_budget_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    want = np.array([sum((r + 1) * (k + 1) for r in range(world)) for k in range(8)], float)
    want[7] = 24.0
    for r in range(world):
        assert np.array_equal(np.load(os.path.join(tmp_path, f"b{r}.npy")), want)


def test_tile_neighbours_match_the_process_grid():
    import noahmp_b200
    from noahmp_b200 import halo
    for world in (1, 2, 3, 4, 6, 8, 12, 16):
        npx, npy = noahmp_b200.proc_grid(world)
        for r in range(world):
            nb = noahmp_b200.tile_neighbours(world, r)
            assert nb == tuple(-1 if x is None else x for x in halo.neighbours(r, world))
            ipx, ipy = r % npx, r // npx
            assert (nb[0] >= 0) == (ipx > 0) and (nb[1] >= 0) == (ipx < npx - 1)
            assert (nb[2] >= 0) == (ipy > 0) and (nb[3] >= 0) == (ipy < npy - 1)
            if nb[1] >= 0:  # neighbours in a process row share the tile height; in a process column, the width
                a, b = noahmp_b200.tile(101, 67, world, r), noahmp_b200.tile(101, 67, world, nb[1])
                assert (a[2], a[3]) == (b[2], b[3]) and b[0] == a[1] + 1
            if nb[3] >= 0:
                a, b = noahmp_b200.tile(101, 67, world, r), noahmp_b200.tile(101, 67, world, nb[3])
                assert (a[0], a[1]) == (b[0], b[1]) and b[2] == a[3] + 1
