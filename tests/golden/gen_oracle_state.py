"""Regression fixture of the CPU oracle itself: CRC-32 of every INOUT/OUT array after 24 steps of configuration C1
(10 x 10, the reference's own CPU-runnable case) and 6 steps of a 48 x 32 window of C4 (land / glacier / water mix),
in the portable-math mode, whose arithmetic does not depend on the host libm.  Not reference output (that is
reference_vectors.npz, beside this file): it pins the oracle against accidental change between rounds.
usage: python tests/golden/gen_oracle_state.py   (rewrites tests/golden/oracle_state.json)"""
import json
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from noahmp_b200 import _capi, synthetic as S, tables  # noqa: E402
from helpers import make_case, run_oracle  # noqa: E402

CASES = [("C1", 10, 10, 24), ("C4", 48, 32, 6)]


def state_crcs():
    td = tables.default_tables("USGS")
    ts = _capi.tables_from_dict(td)
    out = {}
    for name, ni, nj, steps in CASES:
        cfg = S.named_config(name)
        cfg.ni, cfg.nj = ni, nj
        _, st, state = make_case(cfg, td)
        assert run_oracle(cfg, ts, st, state, steps, math_mode=1, nthreads=2) is None
        out[f"{name}_{ni}x{nj}_{steps}steps"] = {n: zlib.crc32(np.ascontiguousarray(state[n]).tobytes())
                                               for n in _capi.INOUT_NAMES + _capi.OUT_NAMES}
    return out


if __name__ == "__main__":
    with open(os.path.join(ROOT, "tests", "golden", "oracle_state.json"), "w") as f:
        json.dump(state_crcs(), f, indent=1, sort_keys=True)
    print("written")
