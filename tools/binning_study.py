"""Offline study (CPU oracle): how well do sort keys for the physical column re-binning predict the canopy Newton
iteration count of LATER steps?  Reports the lane efficiency of the VEGE_FLUX loop, sum(iters) / sum_warps(32 * max
iters in the warp), for several keys and key ages.  usage: python tools/binning_study.py [ni nj nsteps]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from noahmp_b200 import _capi, synthetic as S, tables
from oracle import oracle as O

ni, nj, nsteps = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (384, 256, 30)))
cfg = S.named_config("C3"); cfg.ni, cfg.nj = 4608, 3840
td = tables.default_tables("USGS"); ts = _capi.tables_from_dict(td)
xp = S.backend()
st = S.static_fields(xp, cfg, 1000, 1000 + ni - 1, 1500, 1500 + nj - 1)
state = S.cold_start(cfg, st, S.forcing(xp, cfg, 1, st), td)
O.build()
iters, snow = [], []
for step in range(1, nsteps + 1):
    frc = S.forcing(xp, cfg, step, st)
    arr, sc = S.args_from(cfg, st, frc, state, step)
    sc.update(ims=1000, ime=1000 + ni - 1, its=1000, ite=1000 + ni - 1, jms=1500, jme=1500 + nj - 1, jts=1500, jte=1500 + nj - 1,
              ide=cfg.ni, jde=cfg.nj)
    s, it = O.noahmplsm(arr, sc, ts, nthreads=os.cpu_count(), want_iters=True)
    assert s.code == 0, s.code
    iters.append(it.ravel().copy()); snow.append((-state["isnowxy"]).ravel().copy())
iters = np.array(iters); snow = np.array(snow)

def eff(order, it):
    x = it[order]
    n = (len(x) // 32) * 32
    w = x[:n].reshape(-1, 32)
    return w.sum() / (32.0 * w.max(axis=1).sum())

def bucket5(p):
    return np.where(p == 0, 0, np.where(p <= 6, 1, np.where(p <= 8, 2, np.where(p <= 12, 3, 4))))

print("iteration histogram at the last step:", np.bincount(iters[-1], minlength=21).tolist())
print("mean iters", iters[-1].mean())
base = 8
for age in (1, 2, 5, 10, 19):
    t = base + age
    if t >= nsteps: break
    prev, sn = iters[base], snow[base]
    keys = {
        "grid order": np.zeros_like(prev),
        "bucket5*4+snow (shipped)": bucket5(prev) * 4 + sn,
        "exact count*4+snow": prev * 4 + sn,
        "exact count": prev,
        "veg/noveg*4+snow": (prev > 0) * 4 + sn,
        "max of last 2 steps": np.maximum(iters[base], iters[base - 1]) * 4 + sn,
        "veg/noveg, snow, VEGTYP": ((prev > 0) * 4 + sn) * 32 + st["ivgtyp"].ravel(),
        "veg/noveg, VEGTYP, snow": ((prev > 0) * 32 + st["ivgtyp"].ravel()) * 4 + sn,
        "bucket5, snow, VEGTYP": (bucket5(prev) * 4 + sn) * 32 + st["ivgtyp"].ravel(),
        "veg/noveg, snow, SOILTYP": ((prev > 0) * 4 + sn) * 32 + st["isltyp"].ravel(),
    }
    print(f"key from step {base+1}, evaluated on step {t+1} (age {age}):")
    for name, k in keys.items():
        order = np.argsort(k, kind="stable")
        print(f"   {name:28s} lane efficiency {eff(order, iters[t]):.3f}")
print("oracle (sorted by the step's own count):", eff(np.argsort(iters[-1], kind='stable'), iters[-1]))
