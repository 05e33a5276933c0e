"""BASELINE config 5: CONUS 1 km with opt_run=5 (Miguez-Macho & Fan groundwater).  Times, per model step, the column
physics (device-resident forcing) and the WTABLE_mmf_noahmp call that follows it every step (WTDDT = 30 min, DT = 1 h
-> STEPWTD = 1), with the KCELL/HEAD halo exchange over NCCL when launched with torchrun on N > 1 GPUs.
usage: python tools/c5_groundwater_times.py [steps]     |     torchrun --nproc-per-node N tools/c5_groundwater_times.py"""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import noahmp_b200
from noahmp_b200 import halo, synthetic as S, tables

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
cfg = S.named_config("C3"); cfg.opts["iopt_run"] = 5
td = tables.default_tables("USGS")
xs, xe, ys, ye = noahmp_b200.tile(cfg.ni, cfg.nj, world, rank)
ni, nj = xe - xs + 1, ye - ys + 1
xp = S.backend()
st = S.static_fields(xp, cfg, xs, xe, ys, ye)
frc1 = S.forcing(xp, cfg, 1, st)
state = S.cold_start(cfg, st, frc1, td)
wt, wsc = S.groundwater_fields(cfg, st, state)
bounds = dict(ims=xs, ime=xe, its=xs, ite=xe, jms=ys, jme=ye, jts=ys, jte=ye, ide=cfg.ni, jde=cfg.nj)
wsc.update(bounds)
m = noahmp_b200.NoahMP(td, ni, nj, device=local, sync=noahmp_b200.SYNC_RESIDENT)
arr, sc = S.args_from(cfg, st, frc1, state, 1); sc.update(bounds)
m.upload(arr, sc)
xt = S.backend(dev); st_t = S.static_fields(xt, cfg, xs, xe, ys, ye)
order = ["coszin", "t", "qv", "u", "v", "swdown", "glw", "p", "p", "rainbl", "vegfra", "dz8w"]
ring = []
for h in range(4):
    f = S.forcing(xt, cfg, 1 + h, st_t); pl = {k: f[k].contiguous() for k in set(order) - {"vegfra", "dz8w"}}
    pl["vegfra"] = st_t["vegfra"].contiguous(); pl["dz8w"] = torch.full((nj, ni), 60.0, device=dev); ring.append([pl[k] for k in order])
stream = torch.cuda.Stream(device=dev)

def one(k):
    yr, jul, _ = S.clock(cfg, 1 + k)
    m.bind_forcing([t.data_ptr() for t in ring[k % 4]])
    m.step_device(1 + k, yr, float(jul), float(cfg.dt), stream.cuda_stream)
    stream.synchronize()
    t0 = time.perf_counter()
    m.wtable_begin(wt, wsc)
    if world > 1:
        kc, hd = m.wtable_halo()
        halo.exchange_halo(torch.as_tensor(kc, device=dev), torch.as_tensor(hd, device=dev), rank, world)
        torch.cuda.synchronize()
    m.wtable_end(wt, wsc)
    torch.cuda.synchronize()
    return time.perf_counter() - t0

for k in range(4):
    one(k)
if world > 1:
    dist.barrier()
torch.cuda.synchronize(); t0 = time.perf_counter(); tw = 0.0
for k in range(4, 4 + steps):
    tw += one(k)
torch.cuda.synchronize(); total = time.perf_counter() - t0
v = torch.tensor([total, tw], device=dev)
if world > 1:
    dist.all_reduce(v, op=dist.ReduceOp.MAX)
if rank == 0:
    s = m.status()
    print(json.dumps({"config": "C5: CONUS 1 km, dveg=2, opt_run=5, WTABLE every step", "n_gpus": world, "steps": steps,
                      "ms_per_step_total": 1e3 * float(v[0]) / steps, "ms_per_step_wtable": 1e3 * float(v[1]) / steps,
                      "column_steps_per_s": cfg.ni * cfg.nj * steps / float(v[0]), "status": s.code,
                      "halo": "NCCL p2p, 2 phases" if world > 1 else "none (single tile)"}))
m.close()
if world > 1:
    dist.destroy_process_group()
