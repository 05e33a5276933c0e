"""Per-step device times of the RESIDENT-mode physics on a quarter of CONUS (2304 x 1920, 4.4 M columns), with the
land columns re-binned every `interval` steps.  The forcing cycles through a ring of 4 hours resident in HBM.
usage: python tools/step_times.py INTERVAL [NSTEPS]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.getcwd())
import noahmp_b200
from noahmp_b200 import synthetic as S, tables

interval = int(sys.argv[1])
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
ni, nj = 2304, 1920
cfg = S.named_config("C3")
cfg.ni, cfg.nj = ni, nj
td = tables.default_tables("USGS")
xp = S.backend()
st = S.static_fields(xp, cfg)
frc1 = S.forcing(xp, cfg, 1, st)
state = S.cold_start(cfg, st, frc1, td)
m = noahmp_b200.NoahMP(td, ni, nj, sync=noahmp_b200.SYNC_RESIDENT)
m.set_rebin(interval)
arr, sc = S.args_from(cfg, st, frc1, state, 1)
m.upload(arr, sc)

dev = torch.device("cuda", 0)
xt = S.backend(dev)
st_t = S.static_fields(xt, cfg)
order = ["coszin", "t", "qv", "u", "v", "swdown", "glw", "p", "p", "rainbl", "vegfra", "dz8w"]
ring = []
for h in range(4):
    f = S.forcing(xt, cfg, 1 + h, st_t)
    pl = {k: f[k].contiguous() for k in set(order) - {"vegfra", "dz8w"}}
    pl["vegfra"] = st_t["vegfra"].contiguous()
    pl["dz8w"] = torch.full((nj, ni), 60.0, device=dev)
    ring.append([pl[k] for k in order])
torch.cuda.synchronize()
stream = torch.cuda.Stream(device=dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(nsteps + 1)]
ev[0].record(stream)
for k in range(nsteps):
    yr, jul, _ = S.clock(cfg, 1 + k)
    m.bind_forcing([t.data_ptr() for t in ring[k % 4]])
    m.step_device(1 + k, yr, float(jul), 3600.0, stream.cuda_stream)
    ev[k + 1].record(stream)
torch.cuda.synchronize()
tt = [ev[k].elapsed_time(ev[k + 1]) for k in range(nsteps)]
print("interval", interval, "rebins", m.rebins,
      "mean %.3f (after step 4: %.3f)" % (sum(tt) / nsteps, sum(tt[4:]) / (nsteps - 4)), " ".join("%.2f" % t for t in tt))
s = m.status()
print("status code", s.code, "count", s.count)
for f in ("tsk", "sfcrunoff", "zwtxy", "hfx"):
    m.fetch(arr, sc, f)
    a = np.asarray(arr[f], dtype=np.float64)
    print(f, "mean %.6f min %.4f max %.4f nan %d" % (np.nanmean(a), np.nanmin(a), np.nanmax(a), int(np.isnan(a).sum())))
