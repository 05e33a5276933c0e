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
