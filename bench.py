#!/usr/bin/env python
"""bench.py — column-timesteps/sec of the Noah-MP column-physics step (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # the CUDA path (N>1: under torchrun)
    python bench.py --impl reference --gpus N --steps K ...    # the CPU arm: the C++ oracle on the host cores

A "step" is one pass of the hot path (one `noahmplsm` call = one hourly model step) over the whole CONUS 1 km
domain (4608x3840, dveg=2, 40 % of columns with a 3-layer snow pack: BASELINE.json configs[2], the configuration
the metric is quoted on; it fits one B200).  With N GPUs the domain is tiled exactly as
mpp/module_mpp_land.F90 does (strong scaling of the fixed domain, no communication on the step path).

  value : column-steps/s with state AND forcing already resident in HBM (device-timed, CUDA events, max over ranks)
  e2e   : the same metric through the reference-facing C-ABI call noahmp_b200_noahmplsm() with HOST forcing
          buffers (RESIDENT state mode): every step copies that hour's 12 forcing planes host->device and reads
          TSK/HFX/LH/GRDFLX back to the host arrays.
  roofline : HBM roofline of the dominant kernel (land_kernel): 824 algorithmic bytes per column-step
           / its CUDA-event duration, against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline : the C++ oracle ("port": the Fortran reference cannot be compiled in this image) on the host
          cores, on a bounded sample of the same workload.
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
# from the ncu --set full capture of the CONUS launch (profiles/r01_ncu_land_conus_v10_summary.txt): DRAM bytes moved and
# warp-instructions executed per column by land_kernel<dynveg>
NCU_DRAM_BYTES_PER_COLUMN = (7.381587e9 + 8.687775e9) / 17694720   # profiles/r01_ncu_land_conus_v10_summary.txt
NCU_WARP_INSTR_PER_COLUMN = 4832070937 / 17694720
N_SM, SCHED_PER_SM = 148, 4
FORCING_ORDER = ["coszin", "t", "qv", "u", "v", "swdown", "glw", "p", "p", "rainbl", "vegfra", "dz8w"]


PCIE_H2D_GBPS = 55.5  # PCIe 5 x16 of the B200 box, tools/pcie_probe.py


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C3")
    ap.add_argument("--grid", type=int, nargs=2, default=None, help="override ni nj (testing only)")
    ap.add_argument("--cpu-sample", type=int, nargs=3, default=[1536, 1280, 40], help="ni nj steps of the CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--chunks", type=int, default=0, help="row chunks of the e2e pipeline (0 = library default)")
    ap.add_argument("--math", default="fast", choices=["fast", "parity"])
    return ap.parse_args()


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
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


def cpu_oracle_sample(cfg, tables_dict, ni, nj, nsteps, threads):
    """Time the C++ oracle (host libm, `threads` host threads) on an ni x nj window of the workload."""
    from noahmp_b200 import _capi, synthetic as S
    from oracle import oracle as O
    O.build()
    ts = _capi.tables_from_dict(tables_dict)
    xp = S.backend()
    x0, y0 = (cfg.ni - ni) // 2 + 1, (cfg.nj - nj) // 2 + 1
    st = S.static_fields(xp, cfg, x0, x0 + ni - 1, y0, y0 + nj - 1)
    state = S.cold_start(cfg, st, S.forcing(xp, cfg, 1, st), tables_dict)
    ncol = int((st["xland"] < 1.5).sum())
    O.set_math_mode(0)
    elapsed = 0.0
    for step in range(1, nsteps + 1):
        arr, sc = S.args_from(cfg, st, S.forcing(xp, cfg, step, st), state, step)
        t0 = time.perf_counter()
        status, _ = O.noahmplsm(arr, sc, ts, nthreads=threads)
        elapsed += time.perf_counter() - t0
        if status.code:
            raise RuntimeError(f"oracle conservation check failed at step {step}: {status.code}")
    return ncol * nsteps / elapsed, ncol, elapsed


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The Fortran cannot be built here (no
    Fortran compiler, SURVEY.md finding 1), so this is the line-by-line C++ oracle port on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from noahmp_b200 import synthetic as S, tables
    cfg = S.named_config(args.config)
    if args.grid:
        cfg.ni, cfg.nj = args.grid
    td = tables.default_tables("USGS")
    threads = os.cpu_count() or 1
    ni, nj, _ = args.cpu_sample
    ni, nj = min(ni, cfg.ni), min(nj, cfg.nj)
    total = args.steps + args.warmup
    # each "step" = one hourly step of the ni x nj sample window
    from noahmp_b200 import _capi
    from oracle import oracle as O
    O.build()
    ts = _capi.tables_from_dict(td)
    xp = S.backend()
    x0, y0 = (cfg.ni - ni) // 2 + 1, (cfg.nj - nj) // 2 + 1
    st = S.static_fields(xp, cfg, x0, x0 + ni - 1, y0, y0 + nj - 1)
    state = S.cold_start(cfg, st, S.forcing(xp, cfg, 1, st), td)
    ncol = int((st["xland"] < 1.5).sum())
    O.set_math_mode(0)
    elapsed = 0.0
    for step in range(1, total + 1):
        arr, sc = S.args_from(cfg, st, S.forcing(xp, cfg, step, st), state, step)
        t0 = time.perf_counter()
        O.noahmplsm(arr, sc, ts, nthreads=threads)
        dt = time.perf_counter() - t0
        if step > args.warmup:
            elapsed += dt
    v = ncol * args.steps / elapsed
    sample = f"{ni}x{nj} window ({ncol} columns) of {cfg.name} {cfg.ni}x{cfg.nj}, {args.steps} hourly steps"
    print(json.dumps({
        "impl": "reference", "metric": "column-timesteps/sec", "value": v, "unit": "column-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(cfg, args.gpus),
        "cpu_baseline": {"value": v, "unit": "column-steps/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "column-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def workload_config(cfg, n):
    title = {"C3": "CONUS 1 km", "C4": "global 0.05 deg land mask incl. glacier", "C2": "NLDAS 0.125 deg",
             "C1": "HRLDAS 10x10"}.get(cfg.name, cfg.name)
    return {"workload": f"{cfg.name}: {title} {cfg.ni}x{cfg.nj} hourly NoahMP step, dveg={cfg.opts['idveg']} "
                        f"opt_run={cfg.opts['iopt_run']}, {int(cfg.snow_frac * 100)}% columns with 3-layer snow",
            "grid": [cfg.ni, cfg.nj], "tiling": f"mpp_land_partition {n} rank(s)",
            "l2": "inputs larger than L2 (state+forcing per rank >> 126 MB); no flush needed",
            "parallelism": f"domain tiles x{n}, no collective on the step path"}


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
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = S.named_config(args.config)
    if args.grid:
        cfg.ni, cfg.nj = args.grid
    td = tables.default_tables("USGS")
    xs, xe, ys, ye = noahmp_b200.tile(cfg.ni, cfg.nj, world, rank)
    ni, nj = xe - xs + 1, ye - ys + 1

    # ---- initial state (cold start on the device) and upload: not timed ------------------------------------------------
    xp = S.backend()
    st = S.static_fields(xp, cfg, xs, xe, ys, ye)
    frc1 = S.forcing(xp, cfg, 1, st)
    math = noahmp_b200.MATH_PARITY if args.math == "parity" else noahmp_b200.MATH_FAST
    model = noahmp_b200.NoahMP(td, ni, nj, device=local, sync=noahmp_b200.SYNC_RESIDENT, math=math)
    state = S.cold_start_device(model, cfg, st, frc1, xs, ys)  # NOAHMP_INIT through the library (row f1)
    if args.chunks:
        model.set_chunks(args.chunks)
    arr, sc = S.args_from(cfg, st, frc1, state, 1)
    sc.update(ims=xs, ime=xe, its=xs, ite=xe, jms=ys, jme=ye, jts=ys, jte=ye, ide=cfg.ni, jde=cfg.nj)
    model.upload(arr, sc)
    census = model.census()
    ncol = census["land"] + census["glacier"]

    # ---- forcing hours resident in HBM (ring of R hours generated on the device) -------------------------
    R = 4
    xt = S.backend(dev)
    st_t = S.static_fields(xt, cfg, xs, xe, ys, ye)
    ring, clocks_yr = [], []
    for h in range(R):
        f = S.forcing(xt, cfg, 1 + h, st_t)
        planes = {k: f[k].contiguous() for k in set(FORCING_ORDER) - {"vegfra", "dz8w"}}
        planes["vegfra"] = st_t["vegfra"].contiguous()
        planes["dz8w"] = torch.full((nj, ni), 60.0, device=dev)
        ring.append([planes[k] for k in FORCING_ORDER])
    torch.cuda.synchronize()
    stream = torch.cuda.Stream(device=dev)  # the physics kernels are launched on this stream and timed on it

    def device_step(k):  # k = 0-based global step counter
        yr, julian, _ = S.clock(cfg, 1 + k)
        model.bind_forcing([t.data_ptr() for t in ring[k % R]])
        model.step_device(1 + k, yr, float(julian), float(cfg.dt), stream.cuda_stream)

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
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev[0].record(stream)
    for s in range(args.steps):
        device_step(k); k += 1
        ev[s + 1].record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    gpu_launches = model.launches - launches0
    step_ms = [ev[s].elapsed_time(ev[s + 1]) for s in range(args.steps)]
    total_ms = ev[0].elapsed_time(ev[-1])
    st1 = model.status()
    if st1.code:
        raise SystemExit(f"model conservation check failed: code {st1.code} at ({st1.i},{st1.j}) value {st1.value}")
    model.bind_forcing(None)

    # ---- end-to-end through the reference-facing call with host forcing ("e2e") ---------------------------
    e2e = None
    if not args.no_e2e:
        host_ring = []
        for h in range(R):
            hf = {}
            f = ring[h]
            names = ["coszin", "t3d", "qv3d", "u_phy", "v_phy", "swdown", "glw", "p8w3d", None, "rainbl", "vegfra",
                     "dz8w"]
            for idx, n in enumerate(names):
                if n is None:
                    continue
                if n in _capi.ATM3D:
                    t = torch.empty((nj, 2, ni), dtype=torch.float32).pin_memory()
                    t[:, 0, :] = f[idx].cpu(); t[:, 1, :] = f[idx].cpu()
                else:
                    t = f[idx].cpu().pin_memory()
                hf[n] = t
            host_ring.append(hf)
        out_names = ["tsk", "hfx", "lh", "grdflx"]
        pinned_out = {n: torch.from_numpy(state[n]).pin_memory() for n in out_names}
        e_arr = dict(arr)
        for n in out_names:
            e_arr[n] = pinned_out[n].numpy()
        model.set_fetch(out_names)
        h2d = sum(4 * ni * nj for _ in range(12))
        d2h = 4 * ni * nj * len(out_names)

        def e2e_step(k):
            yr, julian, _ = S.clock(cfg, 1 + k)
            a = dict(e_arr)
            for n, t in host_ring[k % R].items():
                a[n] = t.numpy()
            s2 = dict(sc)
            s2.update(itimestep=1 + k, yr=yr, julian=float(julian))
            return model.noahmplsm(a, s2)  # forcing upload | physics | TSK/HFX/LH/GRDFLX download, pipelined by row chunks

        for _ in range(max(1, args.warmup)):
            e2e_step(k); k += 1
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            stt = e2e_step(k); k += 1
        barrier()
        e2e_s = time.perf_counter() - t0
        if stt.code:
            raise SystemExit(f"e2e: model check failed code {stt.code}")
        e2e = (e2e_s, h2d, d2h)

    # ---- e2e through the on-device forcing pipeline (row f2): forcing FILES every 3 h, interpolation on the GPU ------
    e2e_f2 = None
    if not args.no_e2e:
        idx = {n: i for i, n in enumerate(FORCING_ORDER)}
        files = []
        for h in range(R):
            f = ring[h]
            d = {"t": f[idx["t"]], "q": f[idx["qv"]], "u": f[idx["u"]], "v": f[idx["v"]], "p": f[7], "lw": f[idx["glw"]],
                 "sw": f[idx["swdown"]], "pcp": f[idx["rainbl"]] / float(cfg.dt), "fpar": f[idx["vegfra"]] / 100.0}
            files.append({n: t.float().cpu().pin_memory().numpy() for n, t in d.items()})
        model.forcing_static(st["xlatin"], st["xlong"], 30.0)
        model.forcing_upload(0, files[0]); model.forcing_upload(1, files[1])
        nfile = 1

        def f2_step(k, count):
            nonlocal nfile
            sub = count % 3
            if sub == 0 and count > 0:  # the model time reached file B
                model.forcing_swap()
                nfile += 1
                model.forcing_upload(1, files[nfile % R])  # asynchronous; overlaps the steps below
            yr, julian, hour = S.clock(cfg, 1 + k)
            jul = model.forcing_apply(float(np.float32(3 - sub) / np.float32(3)), int(julian), int(hour), 0, 0, float(cfg.dt))
            s2 = dict(sc)
            s2.update(itimestep=1 + k, yr=yr, julian=jul)
            return model.noahmplsm_device_forcing(e_arr, s2)

        cnt = 0
        for _ in range(3):
            f2_step(k, cnt); k += 1; cnt += 1
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            stt = f2_step(k, cnt); k += 1; cnt += 1
        barrier()
        e2e_f2 = time.perf_counter() - t0
        if stt.code:
            raise SystemExit(f"e2e (forcing pipeline): model check failed code {stt.code}")

    # ---- reduce over ranks ---------------------------------------------------------------------------------
    vals = torch.tensor([total_ms, float(ncol), e2e[0] if e2e else 0.0, float(gpu_launches), e2e_f2 or 0.0], device=dev,
                        dtype=torch.float64)
    if world > 1:
        mx = vals.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = vals.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        total_ms, e2e_s_max, e2e_f2 = float(mx[0]), float(mx[2]), float(mx[4])
        ncol_all, launches_all = float(sm[1]), int(sm[3])
    else:
        e2e_s_max, ncol_all, launches_all = (e2e[0] if e2e else 0.0), float(ncol), int(gpu_launches)

    if rank == 0:
        peak, peak_src = load_peaks()
        value = ncol_all * args.steps / (total_ms * 1e-3)
        # dominant kernel = land_kernel: the timed region is (memsets +) land kernel per step on this rank
        mean_ms = float(np.mean(step_ms))
        achieved = ALG_BYTES_PER_COLUMN_STEP * ncol / (mean_ms * 1e-3) / 1e9
        try:  # measured FP32 issue peak of this GPU type (thread-instructions/s -> warp-instructions/s)
            with open(os.path.join(ROOT, "profiles", "r01_peaks.json")) as f:
                issue_peak, measured_issue = json.load(f)["ffma_thread_instr_per_s"] / 32.0, True
        except Exception:
            issue_peak, measured_issue = N_SM * SCHED_PER_SM * ((clocks or {}).get("sm_mhz") or 1965.0) * 1e6, False
        line = {
            "metric": "column-timesteps/sec", "value": value, "unit": "column-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(cfg, world),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": NCU_DRAM_BYTES_PER_COLUMN * ncol if cfg.name == "C3" else None,
                         "traffic_source": "ncu dram__bytes_read+write of the CONUS launch, profiles/r01_ncu_land_conus_v10_summary.txt",
                         "peak_source": peak_src, "kernel": f"land_kernel<{model.variant}>",
                         "algorithmic_bytes_per_column_step": ALG_BYTES_PER_COLUMN_STEP,
                         "columns_per_launch": ncol, "kernel_ms": mean_ms,
                         "note": "step = memset x2 + land_kernel (+glacier/sea-ice kernels when present); "
                                 "the physics is FP32/SFU-issue bound, see DESIGN.md"},
            # the binding resource: warp-instruction issue slots (FP32 / SFU pipes), not HBM
            "issue_roofline": {"bound": "fp32/sfu issue", "warp_instr_per_column": NCU_WARP_INSTR_PER_COLUMN,
                               "achieved_warp_instr_per_s": NCU_WARP_INSTR_PER_COLUMN * ncol / (mean_ms * 1e-3),
                               "peak_warp_instr_per_s": issue_peak,
                               "peak_source": "measured FFMA issue rate, tools/peaks.cu -> profiles/r01_peaks.json"
                                              if measured_issue else "nominal 148 SM x 4 schedulers x SM clock",
                               "frac": NCU_WARP_INSTR_PER_COLUMN * ncol / (mean_ms * 1e-3) / issue_peak,
                               "simt_efficiency": 0.772,
                               "source": "instruction count from ncu (profiles/), time and clock measured live"},
            "clocks": clocks, "gpu_launches": launches_all, "census": census, "math": args.math,
        }
        if e2e:
            line["e2e"] = {"value": ncol_all * args.steps / e2e_s_max, "unit": "column-steps/s",
                           "h2d_bytes_per_step": e2e[1], "d2h_bytes_per_step": e2e[2],
                           "call": "noahmp_b200_noahmplsm (RESIDENT state, set_fetch=tsk,hfx,lh,grdflx; 9 row chunks, half-height first and last), pinned host buffers",
                           "ms_per_step": 1e3 * e2e_s_max / args.steps,
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
            sni, snj, sst = args.cpu_sample
            sni, snj = min(sni, cfg.ni), min(snj, cfg.nj)
            threads = os.cpu_count() or 1
            v, nc, el = cpu_oracle_sample(cfg, td, sni, snj, sst, threads)
            line["cpu_baseline"] = {"value": v, "unit": "column-steps/s", "cores": threads, "kind": "port",
                                    "sample": f"{sni}x{snj} window ({nc} columns) x {sst} steps of {cfg.name}, "
                                              f"{el:.1f} s, C++ oracle -O2 host libm"}
        print(json.dumps(line))
    model.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
