"""Where does the time of the host-buffer call (bench.py `e2e`) go?  Times noahmp_b200_noahmplsm in RESIDENT mode on
the CONUS tile of one rank for several pipeline settings: with / without the fetch list, with / without the forcing
hints, different row-chunk counts, pinned memory from cudaHostRegister (numpy) vs cudaHostAlloc (torch).
usage: python tools/e2e_probe.py [ni nj]        (NOAHMP_B200_TRACE=1 prints the per-chunk timeline of every call)
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import noahmp_b200  # noqa: E402
from noahmp_b200 import synthetic as S, tables  # noqa: E402

ni, nj = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4608, 3840)
cfg = S.named_config("C3")
cfg.ni, cfg.nj = ni, nj
td = tables.default_tables("USGS")
dev = torch.device("cuda", 0)
noahmp_b200.bind_numa(0)
xp = S.backend()
st = S.static_fields(xp, cfg)
frc1 = S.forcing(xp, cfg, 1, st)
cudart = torch.cuda.cudart()
HOURS = 4
xt = S.backend(dev)
st_t = S.static_fields(xt, cfg)


REGISTERED = []


def pinned(shape, how):
    if how == "register":
        a = np.empty(shape, np.float32)
        cudart.cudaHostRegister(a.ctypes.data, a.nbytes, 0)
        REGISTERED.append(a)
        return a
    t = torch.empty(shape, dtype=torch.float32, pin_memory=True)
    return t.numpy()


def host_ring(how):
    ring = []
    for h in range(HOURS):
        f = S.forcing(xt, cfg, 13 + h, st_t)  # daytime hours
        hf = {}
        for src, n in {"t": "t3d", "qv": "qv3d", "u": "u_phy", "v": "v_phy", "p": "p8w3d"}.items():
            a = pinned((nj, 2, ni), how)
            torch.from_numpy(a).copy_(torch.stack([f[src], f[src]], dim=1))
            hf[n] = a
        for n in ("coszin", "swdown", "glw", "rainbl"):
            a = pinned((nj, ni), how)
            torch.from_numpy(a).copy_(f[n])
            hf[n] = a
        ring.append(hf)
    return ring


def run(tag, how="register", fetch=("tsk", "hfx", "lh", "grdflx"), hints=True, chunks=0, steps=8):
    model = noahmp_b200.NoahMP(td, ni, nj, device=0, sync=noahmp_b200.SYNC_RESIDENT)
    state = S.cold_start_device(model, cfg, st, frc1)
    arr, sc = S.args_from(cfg, st, frc1, state, 1)
    ring = host_ring(how)
    for n in fetch:
        a = pinned(state[n].shape, how)
        a[...] = state[n]
        arr[n] = a
    model.set_fetch(list(fetch))
    if chunks:
        model.set_chunks(chunks)
    if hints:
        model.set_forcing_hints(7)
    prep = model.prepare(arr, sc)
    for k in range(3):
        model.noahmplsm_prepared(prep, 1 + k, 2017, 15.5, ring[k % HOURS])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(3, 3 + steps):
        s = model.noahmplsm_prepared(prep, 1 + k, 2017, 15.5, ring[k % HOURS])
    torch.cuda.synchronize()
    ms = 1e3 * (time.perf_counter() - t0) / steps
    assert s.code == 0
    nup = 9 if hints else 12
    res = {"tag": tag, "ms_per_call": round(ms, 2), "h2d_MB": round(4e-6 * ni * nj * nup, 1),
           "d2h_MB": round(4e-6 * ni * nj * len(fetch), 1), "chunks": chunks or "auto", "pinned": how}
    print(json.dumps(res), flush=True)
    model.close()
    del ring
    torch.cuda.synchronize()
    while REGISTERED:  # numpy would hand the same addresses to the next run's arrays
        cudart.cudaHostUnregister(REGISTERED.pop().ctypes.data)
    return res


out = [run("default"), run("no fetch", fetch=()), run("no hints", hints=False), run("chunks 5", chunks=5),
       run("chunks 16", chunks=16), run("cudaHostAlloc", how="alloc"), run("cudaHostAlloc no fetch", how="alloc", fetch=())]
path = os.path.join(ROOT, "gpurun_out", "r02_e2e_probe.json")
os.makedirs(os.path.dirname(path), exist_ok=True)
json.dump(out, open(path, "w"), indent=1)
