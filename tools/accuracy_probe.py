"""Difference statistics of a library build against the CPU oracle (host libm) on an NLDAS tile.
usage: python tools/accuracy_probe.py [nsteps] [config] (library chosen with NOAHMP_B200_LIB)"""
import os, sys
import numpy as np
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests"))
import noahmp_b200
from noahmp_b200 import _capi, synthetic as S, tables
from helpers import clone_state, make_case, run_gpu, run_oracle
nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
cfg = S.named_config(sys.argv[2] if len(sys.argv) > 2 else "C3"); cfg.ni, cfg.nj = 232, 112
td = tables.default_tables("USGS"); ts = _capi.tables_from_dict(td)
_, st, state0 = make_case(cfg, td)
a, b = clone_state(state0), clone_state(state0)
m = noahmp_b200.NoahMP(td, cfg.ni, cfg.nj, math=noahmp_b200.MATH_FAST)
e1 = run_oracle(cfg, ts, st, a, nsteps, math_mode=0); e2 = run_gpu(m, cfg, st, b, nsteps)
print("lib", os.environ.get("NOAHMP_B200_LIB", "main"), "variant", m.variant, "errors", e1, e2)
for n in ["tsk", "tgxy", "tvxy", "hfx", "lh", "grdflx", "tslb", "smois", "sh2o", "snow", "snowh", "xlaixy", "chxy", "cmxy",
          "fsaxy", "savxy", "sagxy", "firaxy", "t2mvxy", "rssunxy", "psnxy", "eahxy", "tahxy", "sfcrunoff", "udrunoff"]:
    d = np.abs(a[n].astype(np.float64) - b[n]); d = d[np.isfinite(d)]
    print(f"{n:10s} max {d.max():10.3e}  p99 {np.quantile(d,0.99):10.3e}  mean {d.mean():10.3e}")
print("isnow differs:", (a["isnowxy"] != b["isnowxy"]).mean())
