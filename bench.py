#!/usr/bin/env python
"""bench.py — column-timesteps/sec of the Noah-MP column-physics step (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # the CUDA path (N>1: under torchrun)
    python bench.py --impl reference --gpus N --steps K ...    # the CPU arm: the C++ oracle on the host cores
    python bench.py --config C5 ...                            # opt_run=5: + WTABLE_mmf_noahmp every step (NCCL halo)

A "step" is one pass of the hot path (one `noahmplsm` call = one hourly model step) over the whole CONUS 1 km
domain (4608x3840, dveg=2, 40 % of columns with a 3-layer snow pack: BASELINE.json configs[2], the configuration
the metric is quoted on; it fits one B200).  With N GPUs the domain is tiled exactly as
mpp/module_mpp_land.F90 does (strong scaling of the fixed domain, no communication on the step path).

The forcing is a 24-HOUR DIURNAL CYCLE: model step s (1-based) reads hour (s-1) mod 24 of the cycle that starts at
the configuration's start time (C3: 2017-01-15 00 UTC), so steps W+1 .. W+K of every arm — `value`, `e2e`, the
forcing pipeline, `cpu_baseline` and `--impl reference` — see the same hours of the day; the sunlit fraction of the
timed steps and of the whole day is printed next to each number (night columns skip ALBEDO/TWOSTREAM and STOMATA,
phys/module_sf_noahmplsm.F90:2356, :5389, so the hour of day is part of the workload).

  value : column-steps/s with state AND forcing already resident in HBM (device-timed, CUDA events, max over ranks)
  e2e   : the same metric through the reference-facing C-ABI call noahmp_b200_noahmplsm() with HOST forcing
          buffers (RESIDENT state mode): every step copies that hour's forcing planes host->device and reads
          TSK/HFX/LH/GRDFLX back to the host arrays.
  roofline : HBM roofline of the dominant kernel (land_kernel): 824 algorithmic bytes per column-step
           / its CUDA-event duration, against MEASURED_PEAKS.json hbm_gbs.
  compute_roofline : algorithmic FP32 operations and transcendental calls per column-step counted by the op-counting
          instantiation of the oracle (profiles/r02_opcount.json) / measured FFMA and MUFU issue rates.
  cpu_baseline : the C++ oracle ("port": the Fortran reference cannot be compiled in this image) on the host
          cores, on a bounded sample (every 9th row) of the same workload at the same hours.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALG_BYTES_PER_COLUMN_STEP = 824  # 87 words read + 119 words written at the noahmplsm boundary
ALG_BYTES_WTABLE = 188           # + per land column per WTABLE_mmf_noahmp call (config 5)
N_SM, SCHED_PER_SM = 148, 4
FORCING_ORDER = ["coszin", "t", "qv", "u", "v", "swdown", "glw", "p", "p", "rainbl", "vegfra", "dz8w"]
RING_HOURS = 24
PCIE_H2D_GBPS = 55.5  # PCIe 5 x16 of the B200 box, tools/pcie_probe.py (profiles/r01_pcie.json)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=24)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C3", choices=["C1", "C2", "C3", "C4", "C5"])
    ap.add_argument("--grid", type=int, nargs=2, default=None, help="override ni nj (testing only)")
    ap.add_argument("--cpu-stride", type=int, default=9, help="the CPU arms run every n-th row of the domain")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-port", action="store_true", help="CPU arms: the hand-written oracle even where the translated "
                                                            "reference (oracle/_ref) exists")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--chunks", type=int, default=0, help="row chunks of the e2e pipeline (0 = library default)")
    ap.add_argument("--math", default="fast", choices=["fast", "parity"])
    ap.add_argument("--full-day", action="store_true", help="after the timed region, time 24 more steps (one whole day)")
    return ap.parse_args()


def load_json(*path):
    try:
        with open(os.path.join(ROOT, *path)) as f:
            return json.load(f)
    except Exception:
        return None


def kernel_constants(sunlit):
    """Per-column DRAM bytes and executed warp-instructions of land_kernel<dynveg> from the committed ncu captures of a
    night and a midday CONUS launch, interpolated linearly in the sunlit fraction of the timed steps."""
    d = load_json("profiles", "r02_kernel_constants.json")
    if not d:
        return {}
    n, m, cols = d["night"], d["midday"], float(d["columns"])
    mix = lambda k: (n[k] + sunlit * (m[k] - n[k]))
    return {"dram_bytes_per_column": mix("dram_bytes") / cols, "warp_instr_per_column": mix("warp_instr") / cols,
            "simt_efficiency": mix("not_predicated_off_per_warp") / 32.0, "source": d["source"]}


def load_peaks():
    d = load_json("MEASURED_PEAKS.json")
    if d and "hbm_gbs" in d:
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in self.rows if len(r) >= 6 for k in range(4) if r[2 + k] == "Active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def get_config(args):
    from noahmp_b200 import synthetic as S
    cfg = S.named_config(args.config)
    if args.grid:
        cfg.ni, cfg.nj = args.grid
    return cfg


def ring_step(k):
    """0-based step counter k -> model step whose forcing is read: hour k mod 24 of the first day."""
    return 1 + (k % RING_HOURS)


def cpu_arm(cfg, tables_dict, stride, warm, steps, threads):
    """The C++ oracle (host libm, `threads` host threads) on every `stride`-th row of the domain: `warm` untimed steps,
    then `steps` timed ones — the same step numbers, hence the same hours of the day, as the GPU arm."""
    from noahmp_b200 import _capi, synthetic as S
    from oracle import oracle as O
    O.build()
    ts = _capi.tables_from_dict(tables_dict)
    xp = S.backend()
    stride = max(1, min(stride, cfg.nj))
    st = S.static_fields(xp, cfg, jstride=stride)
    state = S.cold_start(cfg, st, S.forcing(xp, cfg, 1, st), tables_dict)
    if cfg.opts["iopt_run"] == 5:
        wt, wsc = S.groundwater_fields(cfg, st, state)
        wsc.update(ide=st["xland"].shape[1], jde=st["xland"].shape[0])  # the sample is its own domain
    land = st["xland"] < 1.5
    ncol = int(land.sum())
    O.set_math_mode(0)
    elapsed, sunlit = 0.0, []
    for k in range(warm + steps):
        frc = S.forcing(xp, cfg, ring_step(k), st)
        arr, sc = S.args_from(cfg, st, frc, state, 1 + k)
        t0 = time.perf_counter()
        status, _ = O.noahmplsm(arr, sc, ts, nthreads=threads)
        if cfg.opts["iopt_run"] == 5:
            # the row sample has no physical neighbours: the lateral-flow stencil runs on it as on any grid (same work)
            O.wtable(wt, wsc, ts)
        dt = time.perf_counter() - t0
        if status.code:
            raise RuntimeError(f"oracle conservation check failed at step {1 + k}: code {status.code}")
        if k >= warm:
            elapsed += dt
            sunlit.append(float((frc["coszin"][land] > 0).mean()))
    nj = st["xland"].shape[0]
    sample = (f"every {stride}th row of {cfg.name} {cfg.ni}x{cfg.nj} = {cfg.ni}x{nj} cells ({ncol} columns), steps "
              f"{warm + 1}..{warm + steps} of the 24 h cycle (sunlit fraction {np.mean(sunlit):.3f}), {elapsed:.1f} s, "
              f"C++ oracle -O2 host libm")
    return ncol * steps / elapsed, ncol, elapsed, float(np.mean(sunlit)), sample


def _ref_worker(cfg, tables_dict, stride, r0, r1, warm, steps, barrier, queue, so):
    """one process of the translated-reference arm: rows r0..r1-1 of the row sample, stepped warm + steps times; the
    compute of a step starts behind a barrier in every process; -> (columns, per-step seconds, sunlit fractions)"""
    from noahmp_b200 import _capi, synthetic as S
    from oracle.ref import refmodel
    R = refmodel.RefModel(so)
    R.set_tables(_capi.tables_from_dict(tables_dict))
    R.set_math_mode(0)
    xp = S.backend()
    st = S.static_fields(xp, cfg, 1, cfg.ni, 1 + r0 * stride, 1 + (r1 - 1) * stride, jstride=stride)
    state = S.cold_start(cfg, st, S.forcing(xp, cfg, 1, st), tables_dict)
    if cfg.opts["iopt_run"] == 5:
        wt, wsc = S.groundwater_fields(cfg, st, state)
        wsc.update(ide=st["xland"].shape[1], jde=st["xland"].shape[0])
    land = st["xland"] < 1.5
    times, sunlit = [], []
    for k in range(warm + steps):
        frc = S.forcing(xp, cfg, ring_step(k), st)
        arr, sc = S.args_from(cfg, st, frc, state, 1 + k)
        barrier.wait()
        t0 = time.perf_counter()
        R.noahmplsm(arr, sc)
        if cfg.opts["iopt_run"] == 5:
            R.wtable(wt, wsc)
        dt = time.perf_counter() - t0
        if k >= warm:
            times.append(dt)
            sunlit.append(float((frc["coszin"][land] > 0).sum()))
    queue.put((int(land.sum()), times, sunlit))


def reference_arm(cfg, tables_dict, stride, warm, steps, procs, so):
    """The reference's own Fortran text, machine-translated and compiled (oracle/_ref/libnoahmp_ref.so), on every
    `stride`-th row of the domain: `procs` processes (the reference keeps its per-column parameters in module
    variables, so it is parallel by tiles, as under MPI), a step's time = the slowest process."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    stride = max(1, min(stride, cfg.nj))
    nrows = (cfg.nj - 1) // stride + 1
    procs = max(1, min(procs, nrows))
    cuts = [nrows * w // procs for w in range(procs + 1)]
    barrier, queue = ctx.Barrier(procs), ctx.Queue()
    ws = [ctx.Process(target=_ref_worker, args=(cfg, tables_dict, stride, cuts[w], cuts[w + 1], warm, steps, barrier,
                                                queue, so)) for w in range(procs)]
    for w in ws:
        w.start()
    res = [queue.get() for _ in ws]
    for w in ws:
        w.join()
    ncol = sum(r[0] for r in res)
    elapsed = sum(max(r[1][k] for r in res) for k in range(steps))
    sunlit = sum(sum(r[2]) for r in res) / max(1, ncol * steps)
    sample = (f"every {stride}th row of {cfg.name} {cfg.ni}x{cfg.nj} = {cfg.ni}x{nrows} cells ({ncol} columns) in {procs} "
              f"row blocks, steps {warm + 1}..{warm + steps} of the 24 h cycle (sunlit fraction {sunlit:.3f}), {elapsed:.1f} s, "
              f"the reference's Fortran machine-translated to C++ (oracle/ref/f90cxx.py), g++ -O2, host libm, one process "
              f"per block")
    return ncol * steps / elapsed, ncol, elapsed, sunlit, sample, procs


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on all host cores.  No Fortran compiler exists
    here, so the reference's own text runs through its machine translation (oracle/_ref/libnoahmp_ref.so, DESIGN.md
    section 5; `kind: reference`); without that library (never built where no reference tree was present) the
    hand-written C++ oracle stands in (`kind: port`)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from noahmp_b200 import tables
    from oracle.ref import refmodel
    cfg = get_config(args)
    td = tables.default_tables("USGS")
    threads = os.cpu_count() or 1
    so = None if args.cpu_port else refmodel.build()
    if so:
        v, ncol, elapsed, sunlit, sample, threads = reference_arm(cfg, td, args.cpu_stride, args.warmup, args.steps,
                                                                  threads, so)
        kind = "reference"
    else:
        v, ncol, elapsed, sunlit, sample = cpu_arm(cfg, td, args.cpu_stride, args.warmup, args.steps, threads)
        kind = "port"
    print(json.dumps({
        "impl": "reference", "metric": "column-timesteps/sec", "value": v, "unit": "column-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(cfg, args.gpus), "sunlit_fraction": sunlit,
        "cpu_baseline": {"value": v, "unit": "column-steps/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "column-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def workload_config(cfg, n):
    title = {"C3": "CONUS 1 km", "C4": "global 0.05 deg land mask incl. glacier", "C2": "NLDAS 0.125 deg",
             "C1": "HRLDAS 10x10", "C5": "CONUS 1 km, MMF groundwater (WTABLE_mmf_noahmp after every step)"}.get(cfg.name, cfg.name)
    y, mo, d, h = cfg.start
    return {"workload": f"{cfg.name}: {title} {cfg.ni}x{cfg.nj} hourly NoahMP step over a 24 h diurnal cycle of forcing "
                        f"(from {y}-{mo:02d}-{d:02d} {h:02d} UTC), dveg={cfg.opts['idveg']} opt_run={cfg.opts['iopt_run']}, "
                        f"{int(cfg.snow_frac * 100)}% columns with 3-layer snow",
            "grid": [cfg.ni, cfg.nj], "tiling": f"mpp_land_partition {n} rank(s)",
            "forcing": "24 hourly forcing states resident in HBM (value) / in pinned host memory (e2e); timed step s reads "
                       "hour (s-1) mod 24",
            "l2": "inputs larger than L2 (state+forcing per rank >> 126 MB); no flush needed",
            "parallelism": f"domain tiles x{n}, " + ("KCELL/HEAD halo over NCCL inside noahmp_b200_wtable" if cfg.name == "C5"
                                                      else "no collective on the step path")}


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    import torch
    import torch.distributed as dist
    import noahmp_b200
    from noahmp_b200 import _capi, synthetic as S, tables

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: noahmp_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    noahmp_b200.bind_numa(local)  # this rank's threads (and the host buffers they first touch) next to its GPU
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = get_config(args)
    c5 = cfg.opts["iopt_run"] == 5
    td = tables.default_tables("USGS")
    xs, xe, ys, ye = noahmp_b200.tile(cfg.ni, cfg.nj, world, rank)
    ni, nj = xe - xs + 1, ye - ys + 1
    bounds = dict(ims=xs, ime=xe, its=xs, ite=xe, jms=ys, jme=ye, jts=ys, jte=ye, ide=cfg.ni, jde=cfg.nj)

    # ---- initial state (cold start on the device) and upload: not timed ------------------------------------------------
    xp = S.backend()
    st = S.static_fields(xp, cfg, xs, xe, ys, ye)
    frc1 = S.forcing(xp, cfg, 1, st)
    math = noahmp_b200.MATH_PARITY if args.math == "parity" else noahmp_b200.MATH_FAST
    model = noahmp_b200.NoahMP(td, ni, nj, device=local, sync=noahmp_b200.SYNC_RESIDENT, math=math)
    if c5:
        state = S.cold_start(cfg, st, frc1, td)
        wt, wsc = S.groundwater_fields(cfg, st, state)
        wsc.update(bounds)
        if world > 1:  # the library's own NCCL communicator for the KCELL/HEAD halo (and the budget all-reduce)
            uid = torch.zeros(128, dtype=torch.uint8)
            if rank == 0:
                uid = torch.from_numpy(model.comm_unique_id().copy())
            uid = uid.to(dev)
            dist.broadcast(uid, 0)
            model.comm_init(uid.cpu().numpy(), rank, world)
    else:
        state = S.cold_start_device(model, cfg, st, frc1, xs, ys)  # NOAHMP_INIT through the library (row f1)
    if args.chunks:
        model.set_chunks(args.chunks)
    arr, sc = S.args_from(cfg, st, frc1, state, 1)
    sc.update(bounds)
    model.upload(arr, sc)
    census = model.census()
    ncol = census["land"] + census["glacier"]

    # ---- the 24 forcing hours resident in HBM (generated on the device) -------------------------------------------
    xt = S.backend(dev)
    st_t = S.static_fields(xt, cfg, xs, xe, ys, ye)
    land_t = st_t["xland"] < 1.5
    ring, sun = [], []
    dz8w_t = torch.full((nj, ni), 60.0, device=dev)
    vegfra_t = st_t["vegfra"].contiguous()
    for h in range(RING_HOURS):
        f = S.forcing(xt, cfg, 1 + h, st_t)
        planes = {k: f[k].contiguous() for k in set(FORCING_ORDER) - {"vegfra", "dz8w"}}
        planes["vegfra"], planes["dz8w"] = vegfra_t, dz8w_t
        ring.append([planes[k] for k in FORCING_ORDER])
        sun.append(float(((f["coszin"] > 0) & land_t).sum()))
    torch.cuda.synchronize()
    stream = torch.cuda.Stream(device=dev)  # the physics kernels are launched on this stream and timed on it

    def device_step(k):  # k = 0-based global step counter
        yr, julian, _ = S.clock(cfg, ring_step(k))
        model.bind_forcing([t.data_ptr() for t in ring[k % RING_HOURS]])
        model.step_device(1 + k, yr, float(julian), float(cfg.dt), stream.cuda_stream)
        if c5:
            model.wtable_device(wt, wsc, stream.cuda_stream)  # WTABLE_mmf_noahmp incl. the halo exchange, same stream

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ("value") ------------------------------------------------------------------
    k = 0
    for _ in range(args.warmup):
        device_step(k); k += 1
    st0 = model.status()
    if st0.code:
        raise SystemExit(f"model conservation check failed during warm-up: code {st0.code} at ({st0.i},{st0.j})")
    sampler = ClockSampler(local)
    launches0 = model.launches
    barrier()
    if rank == 0:
        sampler.start()
    nsteps = args.steps + (RING_HOURS if args.full_day else 0)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(nsteps + 1)]
    k_first = k
    ev[0].record(stream)
    for s in range(args.steps):
        device_step(k); k += 1
        ev[s + 1].record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    gpu_launches = model.launches - launches0
    ev_day0 = torch.cuda.Event(enable_timing=True)
    ev_day0.record(stream)
    for s in range(args.steps, nsteps):  # optional: one more whole day, outside the K timed steps
        device_step(k); k += 1
        ev[s + 1].record(stream)
    torch.cuda.synchronize()
    step_ms = [(ev_day0 if s == args.steps else ev[s]).elapsed_time(ev[s + 1]) for s in range(nsteps)]
    total_ms = ev[0].elapsed_time(ev[args.steps])
    st1 = model.status()
    if st1.code:
        raise SystemExit(f"model conservation check failed: code {st1.code} at ({st1.i},{st1.j}) value {st1.value}")
    model.bind_forcing(None)
    timed_hours = [(k_first + s) % RING_HOURS for s in range(args.steps)]

    # ---- end-to-end through the reference-facing call with host forcing ("e2e") ---------------------------
    e2e = None
    run_e2e = not args.no_e2e and not c5
    if run_e2e:
        cudart = torch.cuda.cudart()
        names3d = {"t": "t3d", "qv": "qv3d", "u": "u_phy", "v": "v_phy", "p": "p8w3d"}
        names2d = {"coszin": "coszin", "swdown": "swdown", "glw": "glw", "rainbl": "rainbl"}
        idx = {n: i for i, n in enumerate(FORCING_ORDER)}

        def pinned(shape):
            a = np.empty(shape, np.float32)
            cudart.cudaHostRegister(a.ctypes.data, a.nbytes, 0)
            return a

        host_ring = []
        for h in range(RING_HOURS):
            hf = {}
            for src, n in names3d.items():
                a = pinned((nj, 2, ni))
                lev = ring[h][idx[src]]
                torch.from_numpy(a).copy_(torch.stack([lev, lev], dim=1))  # levels 1 and 2 identical (driver :338)
                hf[n] = a
            for src, n in names2d.items():
                a = pinned((nj, ni))
                torch.from_numpy(a).copy_(ring[h][idx[src]])
                hf[n] = a
            host_ring.append(hf)
        out_names = ["tsk", "hfx", "lh", "grdflx"]
        e_arr = dict(arr)
        for n in out_names:
            a = pinned(state[n].shape)
            a[...] = state[n]
            e_arr[n] = a
        model.set_fetch(out_names)
        # DZ8W (= 2*zlvl) never changes, VEGFRA changes when a forcing file brings a new one (not in this workload), and
        # the driver copies level 1 of P8W3D into level 2: declared, those three planes cross PCIe once, not every step
        model.set_forcing_hints(noahmp_b200.HINT_DZ8W_CONSTANT | noahmp_b200.HINT_VEGFRA_UNCHANGED |
                                noahmp_b200.HINT_P8W_LEVELS_EQUAL)
        nup = 9
        h2d = 4 * ni * nj * nup
        d2h = 4 * ni * nj * len(out_names)
        prep = model.prepare(e_arr, sc)

        def e2e_step(k):
            yr, julian, _ = S.clock(cfg, ring_step(k))
            # forcing upload | physics | TSK/HFX/LH/GRDFLX download, pipelined by row chunks
            return model.noahmplsm_prepared(prep, 1 + k, yr, float(julian), host_ring[k % RING_HOURS])

        # same hours of the day as the device-resident timing: resume at the next step whose hour is k_first's
        k += (k_first - args.warmup - k) % RING_HOURS
        for _ in range(max(1, args.warmup)):
            e2e_step(k); k += 1
        barrier()
        t0 = time.perf_counter()
        e2e_calls = []
        for _ in range(args.steps):
            t1 = time.perf_counter()
            stt = e2e_step(k); k += 1
            e2e_calls.append(1e3 * (time.perf_counter() - t1))
        barrier()
        e2e_s = time.perf_counter() - t0
        if stt.code:
            raise SystemExit(f"e2e: model check failed code {stt.code}")
        e2e = (e2e_s, h2d, d2h, nup, e2e_calls)
        model.set_forcing_hints(0)

    # ---- e2e through the on-device forcing pipeline (row f2): forcing FILES every 3 h, interpolation on the GPU ------
    e2e_f2 = None
    if run_e2e:
        files = []
        for h in range(0, RING_HOURS, 3):
            f = ring[h]
            d = {"t": f[idx["t"]], "q": f[idx["qv"]], "u": f[idx["u"]], "v": f[idx["v"]], "p": f[7], "lw": f[idx["glw"]],
                 "sw": f[idx["swdown"]], "pcp": f[idx["rainbl"]] / float(cfg.dt), "fpar": f[idx["vegfra"]] / 100.0}
            hf = {}
            for n, t in d.items():
                a = pinned((nj, ni))
                torch.from_numpy(a).copy_(t.float())
                hf[n] = a
            files.append(hf)
        nf = len(files)
        model.forcing_static(st["xlatin"], st["xlong"], 30.0)
        k += (k_first - 3 - k) % RING_HOURS
        fa = ((k % RING_HOURS) // 3) % nf
        model.forcing_upload(0, files[fa]); model.forcing_upload(1, files[(fa + 1) % nf])
        state_f2 = {"file": fa, "first": True}

        def f2_step(k):
            hour = k % RING_HOURS
            sub = hour % 3
            if sub == 0 and not state_f2["first"]:  # the model time reached file B
                model.forcing_swap()
                state_f2["file"] = (state_f2["file"] + 1) % nf
                model.forcing_upload(1, files[(state_f2["file"] + 1) % nf])  # asynchronous; overlaps the steps below
            state_f2["first"] = False
            yr, julian, hr = S.clock(cfg, ring_step(k))
            jul = model.forcing_apply(float(np.float32(3 - sub) / np.float32(3)), int(julian), int(hr), 0, 0, float(cfg.dt))
            return model.noahmplsm_prepared(prep, 1 + k, yr, jul, None, device_forcing=True)

        for _ in range(3):
            f2_step(k); k += 1
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            stt = f2_step(k); k += 1
        barrier()
        e2e_f2 = time.perf_counter() - t0
        if stt.code:
            raise SystemExit(f"e2e (forcing pipeline): model check failed code {stt.code}")

    # ---- reduce over ranks ---------------------------------------------------------------------------------
    vals = torch.tensor([total_ms, float(ncol), e2e[0] if e2e else 0.0, float(gpu_launches), e2e_f2 or 0.0,
                         float(census["land"])] + sun + step_ms, device=dev, dtype=torch.float64)
    if world > 1:
        mx = vals.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = vals.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        total_ms, e2e_s_max, e2e_f2 = float(mx[0]), float(mx[2]), float(mx[4])
        ncol_all, launches_all, nland_all = float(sm[1]), int(sm[3]), float(sm[5])
        sun_all = [float(x) for x in sm[6:6 + RING_HOURS]]
        step_ms_max = [float(x) for x in mx[6 + RING_HOURS:]]
    else:
        e2e_s_max, ncol_all, launches_all, nland_all = (e2e[0] if e2e else 0.0), float(ncol), int(gpu_launches), float(census["land"])
        sun_all, step_ms_max = sun, step_ms
    if rank == 0:
        peak, peak_src = load_peaks()
        # sunlit fraction of the (non-water) cells, per hour of the cycle
        sun_frac = [s / max(ncol_all, 1.0) for s in sun_all]
        sun_timed = float(np.mean([sun_frac[h] for h in timed_hours]))
        value = ncol_all * args.steps / (total_ms * 1e-3)
        # dominant kernel = land_kernel: the timed region is land kernel (+ re-binning every 20 steps, + glacier / sea-ice
        # kernels when present) per step on this rank
        mean_ms = float(np.mean(step_ms[:args.steps]))
        alg_bytes = ALG_BYTES_PER_COLUMN_STEP + (ALG_BYTES_WTABLE if c5 else 0)
        achieved = alg_bytes * ncol / (mean_ms * 1e-3) / 1e9
        kc = kernel_constants(sun_timed)
        peaks = load_json("profiles", "r01_peaks.json") or {}
        opc = load_json("profiles", "r02_opcount.json") or {}
        line = {
            "metric": "column-timesteps/sec", "value": value, "unit": "column-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(cfg, world),
            "sunlit_fraction": sun_timed, "sunlit_fraction_24h": float(np.mean(sun_frac)),
            "timed_hours_utc": [(cfg.start[3] + h) % 24 for h in timed_hours],
            "step_ms": [round(x, 3) for x in step_ms_max[:args.steps]],
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (kc.get("dram_bytes_per_column") * ncol) if kc.get("dram_bytes_per_column") and cfg.name == "C3" else None,
                         "traffic_source": kc.get("source"),
                         "peak_source": peak_src, "kernel": f"land_kernel<{model.variant}>",
                         "algorithmic_bytes_per_column_step": alg_bytes,
                         "columns_per_launch": ncol, "kernel_ms": mean_ms,
                         "note": "step = land_kernel (+ the re-binning check every 20 steps and its permutation when due; + glacier/sea-ice kernels when "
                                 "present); the physics is FP32/SFU-issue and latency bound, see compute_roofline and DESIGN.md"},
            "clocks": clocks, "gpu_launches": launches_all, "census": census, "math": args.math,
            "rebinning": {"permutations_rank0": int(model.rebins),
                          "policy": "every 20 steps the bin keys are recomputed and the columns permuted only if more than "
                                    "0.1 % of them left their bin (NOAHMP_B200_REBIN_MIN_CHANGED; 0 = always: +0.16 ms per "
                                    "step on this workload, profiles/r02_notes.md)"},
        }
        if args.full_day:
            day = step_ms_max[args.steps:args.steps + RING_HOURS]
            hours = [(k_first + args.steps + s) % RING_HOURS for s in range(RING_HOURS)]
            line["full_day"] = {"ms_per_step": float(np.mean(day)), "value": ncol_all / (float(np.mean(day)) * 1e-3),
                                "by_hour_utc": {str((cfg.start[3] + h) % 24): round(t, 3) for h, t in zip(hours, day)},
                                "sunlit_by_hour_utc": {str((cfg.start[3] + h) % 24): round(sun_frac[h], 3) for h in hours},
                                "note": "24 consecutive steps after the K timed ones, CUDA events, max over ranks"}
        # the binding resource: FP32 / SFU issue. Numerator = ALGORITHMIC operations per column-step counted by the
        # op-counting instantiation of the oracle on this workload (tools/opcount.py), not executed instructions.
        if opc.get(cfg.name if cfg.name != "C5" else "C3") and peaks:
            o = opc[cfg.name if cfg.name != "C5" else "C3"]
            hrs = [o["by_hour_utc"][str((cfg.start[3] + h) % 24)] for h in timed_hours]
            fp32_pc = float(np.mean([x["fp32_instr"] for x in hrs]))
            fp32u_pc = float(np.mean([x["fp32_instr_unfused"] for x in hrs]))
            mufu_pc = float(np.mean([x["mufu"] for x in hrs]))
            colrate = ncol / (mean_ms * 1e-3)
            fpk, mpk = peaks["ffma_thread_instr_per_s"], peaks["mufu_ex2_thread_instr_per_s"]
            f_fp32, f_mufu = fp32_pc * colrate / fpk, mufu_pc * colrate / mpk
            f_issue = (fp32_pc + mufu_pc) * colrate / fpk
            line["compute_roofline"] = {
                "bound": "sfu (MUFU pipe)" if f_mufu >= f_issue else "fp32 issue",
                "frac": max(f_mufu, f_issue),
                "fp32_instr_per_column_step": fp32_pc, "fp32_instr_unfused_per_column_step": fp32u_pc,
                "mufu_per_column_step": mufu_pc,
                "mufu": {"achieved_thread_instr_per_s": mufu_pc * colrate, "peak": mpk, "frac": f_mufu},
                "fp32": {"achieved_thread_instr_per_s": fp32_pc * colrate, "peak": fpk, "frac": f_fp32},
                "issue": {"achieved_thread_instr_per_s": (fp32_pc + mufu_pc) * colrate, "peak": fpk, "frac": f_issue},
                "source": "ALGORITHMIC operations per column-step counted by the op-counting instantiation of the oracle "
                          "on the timed hours (tools/opcount.py -> profiles/r02_opcount.json: " + o.get("sample", "") + "; the "
                          "reference's own text, translated and compiled with the same counting type, gives the same "
                          "transcendental counts and adds / multiplies / divides within 0.7 %: "
                          "profiles/r02_opcount_reference_check.txt); "
                          "fp32_instr assumes every add fuses with a multiply (lower bound), divisions and transcendentals "
                          "expanded as the production build issues them; peaks = measured FFMA and MUFU.EX2 issue rates of "
                          "this GPU type (tools/peaks.cu, profiles/r01_peaks.json); frac = the larger of the MUFU-pipe and "
                          "the issue-slot fraction"}
        if kc.get("warp_instr_per_column"):
            wi = kc["warp_instr_per_column"] * ncol / (mean_ms * 1e-3)
            issue_peak = (peaks.get("ffma_thread_instr_per_s") or N_SM * SCHED_PER_SM * 32 * 1.965e9) / 32.0
            line["issue_utilisation"] = {"executed_warp_instr_per_column": kc["warp_instr_per_column"],
                                         "achieved_warp_instr_per_s": wi, "peak_warp_instr_per_s": issue_peak,
                                         "frac": wi / issue_peak, "simt_efficiency": kc.get("simt_efficiency"),
                                         "source": kc.get("source"),
                                         "note": "executed instructions (a utilisation, not an algorithmic roofline)"}
        if e2e:
            line["e2e"] = {"value": ncol_all * args.steps / e2e_s_max, "unit": "column-steps/s",
                           "h2d_bytes_per_step": e2e[1], "d2h_bytes_per_step": e2e[2],
                           "call": "noahmp_b200_noahmplsm (RESIDENT state, set_fetch=tsk,hfx,lh,grdflx; row-chunk pipeline; "
                                   f"{e2e[3]} forcing planes per call: DZ8W, VEGFRA and level 2 of P8W3D declared "
                                   "constant / unchanged / equal to level 1 with noahmp_b200_set_forcing_hints), pinned host buffers",
                           "ms_per_step": 1e3 * e2e_s_max / args.steps, "sunlit_fraction": sun_timed,
                           "call_ms_rank0": [round(x, 2) for x in e2e[4]],
                           # the forcing upload is what bounds this call: bytes per rank / the box's pinned H2D rate
                           "pcie_floor_ms": e2e[1] / PCIE_H2D_GBPS / 1e6,
                           "pcie_note": f"{PCIE_H2D_GBPS} GB/s pinned H2D measured with tools/pcie_probe.py "
                                        "(profiles/r01_pcie.json)"}
            line["e2e_forcing_pipeline"] = {
                "value": ncol_all * args.steps / e2e_f2, "unit": "column-steps/s", "ms_per_step": 1e3 * e2e_f2 / args.steps,
                "h2d_bytes_per_step": 9 * 4 * ni * nj // 3, "d2h_bytes_per_step": e2e[2],
                "call": "row f2: noahmp_b200_forcing_upload (9 file fields every 3 steps, async) + forcing_apply + "
                        "noahmplsm_device_forcing (same fetch list)"}
        else:
            line["e2e"] = None
        if world == 1 and not args.no_cpu_baseline:
            # the reference arm in a process of its own (it forks workers; this one holds a CUDA context)
            cb = None
            if not args.cpu_port:
                try:
                    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--config", args.config,
                           "--steps", str(args.steps), "--warmup", str(args.warmup), "--cpu-stride", str(args.cpu_stride)]
                    if args.grid:
                        cmd += ["--grid", str(args.grid[0]), str(args.grid[1])]
                    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
                    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env).stdout
                    rl = json.loads([l for l in out.splitlines() if l.startswith("{")][-1])
                    cb = dict(rl["cpu_baseline"], sunlit_fraction=rl["sunlit_fraction"])
                except Exception as e:  # noqa: BLE001  (the baseline is a reported number, not the product)
                    sys.stderr.write(f"reference arm as a subprocess failed ({e}); timing the oracle port\n")
            if cb is None:
                threads = os.cpu_count() or 1
                v, nc, el, sl, sample = cpu_arm(cfg, td, args.cpu_stride, args.warmup, args.steps, threads)
                cb = {"value": v, "unit": "column-steps/s", "cores": threads, "kind": "port", "sample": sample,
                      "sunlit_fraction": sl}
            line["cpu_baseline"] = cb
        print(json.dumps(line))
    model.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
