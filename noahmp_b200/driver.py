"""Host-side driver over the C-ABI: the Python mirror of the reference's `noahmplsm` call.

Arrays use the reference's Fortran memory layout, expressed as C-contiguous numpy arrays of shape
(nj[, k], ni) (see _capi.array_shape); scalars carry the names of the `noahmplsm` dummy arguments.
"""
import collections
import ctypes as C

import numpy as np

from . import _capi, _lib

SYNC_FULL, SYNC_RESIDENT = 0, 1
MATH_FAST, MATH_PARITY = 0, 1
HINT_DZ8W_CONSTANT, HINT_VEGFRA_UNCHANGED, HINT_P8W_LEVELS_EQUAL = 1, 2, 4
HOLD_BUDGET_BYTES = 64 << 30  # arrays kept referenced (and page-locked by the library) per model, least recently used out

# messages the reference passes to wrf_error_fatal for each status code (include/noahmp_b200.h)
ERROR_TEXT = {
    1: "Stop in Noah-MP (ERRSW)", 2: "Energy budget problem in NOAHMP LSM", 3: "Water budget problem in NOAHMP LSM",
    4: "STOP in Noah-MP (emitted longwave <0)", 5: "CRITICAL PROBLEM: HCAN <= ZPD", 6: "STOP in Noah-MP (ZLVL <= ZPD)",
    7: "REDPRM: table index out of range", 8: "unsupported option value",
    9: "lsminit: out of range value of ISLTYP", 100: "CUDA failure", 101: "bad argument",
}


class NoahmpError(RuntimeError):
    def __init__(self, code, detail=""):
        self.code = code
        super().__init__(f"noahmp_b200 error {code}: {ERROR_TEXT.get(code, '?')} {detail}".strip())


def read_tables(directory, dataset="USGS", soil="STAS"):
    """MPTABLE/VEGPARM/SOILPARM/GENPARM.TBL -> _capi.NoahmpTables through the library's C++ reader."""
    t = _capi.NoahmpTables()
    rc = _lib.lib().noahmp_b200_read_tables(str(directory).encode(), dataset.encode(), soil.encode(), C.byref(t))
    if rc:
        raise NoahmpError(rc, _lib.lib().noahmp_b200_tables_error().decode())
    return t


def proc_grid(nproc):
    nx, ny = C.c_int(), C.c_int()
    _lib.lib().noahmp_b200_proc_grid(nproc, C.byref(nx), C.byref(ny))
    return nx.value, ny.value


def tile_neighbours(nproc, rank):
    """(left, right, below, above) ranks of the neighbouring tiles in the mpp_land process grid, -1 at the edge."""
    nb = (C.c_int * 4)()
    _lib.lib().noahmp_b200_tile_neighbours(nproc, rank, nb)
    return tuple(nb)


def tile(global_nx, global_ny, nproc, rank):
    """(xstart, xend, ystart, yend), 1-based inclusive, of `rank` (mpp_land_partition_calc)."""
    v = [C.c_int() for _ in range(4)]
    _lib.lib().noahmp_b200_tile(global_nx, global_ny, nproc, rank, *[C.byref(x) for x in v])
    return tuple(x.value for x in v)


def bind_numa(device):
    """Pin this process's threads to the CPUs of the NUMA node its GPU hangs off (sysfs), so that the pinned host
    buffers it first-touches and the copies it drives stay on that node.  One process per GPU; a no-op when the
    topology cannot be read.  Returns the node or None."""
    import os
    try:
        import torch
        bus = torch.cuda.get_device_properties(device)
        pci = f"{bus.pci_domain_id:04x}:{bus.pci_bus_id:02x}:{bus.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{pci}/numa_node") as f:
            node = int(f.read())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception:
        pass
    return None


class _DevArray:
    """Zero-copy view of a device plane for torch.as_tensor / cupy (CUDA array interface v2)."""

    def __init__(self, ptr, shape, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class NoahMP:
    """One tile of ni x nj grid cells on one GPU."""

    def __init__(self, tables, ni, nj, device=0, sync=SYNC_FULL, math=MATH_FAST):
        self._L = _lib.lib()
        if isinstance(tables, dict):
            tables = _capi.tables_from_dict(tables)
        self.tables = tables
        self.ni, self.nj = ni, nj
        self._ctx = self._L.noahmp_b200_create(device, C.byref(tables), ni, nj)
        if not self._ctx:
            raise NoahmpError(100, self._L.noahmp_b200_last_error().decode())
        self.set_mode(sync)
        self.set_math(math)
        self._held = collections.OrderedDict()
        self._held_bytes = 0

    def _hold(self, arrays):
        """The library page-locks caller arrays of 4 MiB and more and remembers them by address (a Fortran driver's
        arrays live as long as the run).  numpy would unmap a freed array while it is still registered, so arrays of
        that size stay referenced here; a loop that passes fresh arrays to every call is bounded: beyond
        HOLD_BUDGET_BYTES the least recently used ones are un-pinned (noahmp_b200_unpin) and released."""
        for v in (arrays.values() if isinstance(arrays, dict) else arrays):
            if isinstance(v, np.ndarray) and v.nbytes >= (4 << 20):
                key = v.ctypes.data
                if key in self._held:
                    self._held.move_to_end(key)
                else:
                    self._held[key] = v
                    self._held_bytes += v.nbytes
        if self._held_bytes > HOLD_BUDGET_BYTES:
            live = {v.ctypes.data for v in (arrays.values() if isinstance(arrays, dict) else arrays)
                    if isinstance(v, np.ndarray)}
            for key in list(self._held):
                if self._held_bytes <= HOLD_BUDGET_BYTES:
                    break
                if key in live:
                    continue
                self._L.noahmp_b200_unpin(self._ctx, key)
                self._held_bytes -= self._held.pop(key).nbytes
        return arrays

    def close(self):
        if getattr(self, "_ctx", None):
            self._L.noahmp_b200_destroy(self._ctx)
            self._ctx = None
            self._held = collections.OrderedDict()
            self._held_bytes = 0

    __del__ = close

    def _check(self, rc):
        if rc in (100, 101, 8):
            raise NoahmpError(rc, self._L.noahmp_b200_last_error().decode())
        return rc

    def set_mode(self, sync):
        self._check(self._L.noahmp_b200_set_mode(self._ctx, sync))

    def set_math(self, math):
        self._check(self._L.noahmp_b200_set_math(self._ctx, math))

    # ---- the reference-facing call -------------------------------------------------------------
    def noahmplsm(self, arrays, scalars):
        """CALL noahmplsm(...): updates the INOUT/OUT arrays in place (SYNC_FULL) and returns NoahmpStatus."""
        a = _capi.make_args(self._hold(arrays), scalars)
        st = _capi.NoahmpStatus()
        self._check(self._L.noahmp_b200_noahmplsm(self._ctx, C.byref(a), C.byref(st)))
        return st

    def prepare(self, arrays, scalars):
        """Marshal the 158-member argument list once; noahmplsm_prepared() then only patches what changes per step
        (a Fortran caller builds the struct in the shim for free, ctypes needs ~0.5 ms for it)."""
        return [_capi.make_args(self._hold(arrays), scalars), dict(arrays)]

    def noahmplsm_prepared(self, prep, itimestep, yr, julian, forcing=None, device_forcing=False):
        """noahmplsm() with a prepared argument list: `forcing` = dict of the arrays that are different objects this
        step (names of the noahmplsm dummy arguments); device_forcing: the planes filled by forcing_apply() are used
        (noahmplsm_device_forcing)."""
        a, keep = prep
        a.itimestep, a.yr, a.julian = int(itimestep), int(yr), float(julian)
        if forcing:
            self._hold(forcing)
            for n, arr in forcing.items():
                setattr(a, n, arr.ctypes.data_as(_capi.ARG_POINTER_TYPE[n]))
                keep[n] = arr
        st = _capi.NoahmpStatus()
        fn = self._L.noahmp_b200_noahmplsm_device_forcing if device_forcing else self._L.noahmp_b200_noahmplsm
        self._check(fn(self._ctx, C.byref(a), C.byref(st)))
        return st

    def sync_host(self, arrays, scalars):
        a = _capi.make_args(self._hold(arrays), scalars)
        self._check(self._L.noahmp_b200_sync_host(self._ctx, C.byref(a)))

    # ---- device-resident stepping ----------------------------------------------------------------
    def upload(self, arrays, scalars):
        a = _capi.make_args(self._hold(arrays), scalars)
        self._check(self._L.noahmp_b200_upload(self._ctx, C.byref(a)))

    def device_forcing(self):
        """List of 12 device planes (nj, ni) in the order documented in include/noahmp_b200.h."""
        ptrs = (C.c_void_p * 12)()
        self._check(self._L.noahmp_b200_device_forcing(self._ctx, ptrs))
        return [_DevArray(p, (self.nj, self.ni)) for p in ptrs]

    def bind_forcing(self, ptrs=None):
        """ptrs: 12 device addresses (ints) of caller-owned (nj, ni) float32 planes, or None to unbind."""
        if ptrs is None:
            self._check(self._L.noahmp_b200_bind_forcing(self._ctx, None))
        else:
            arr = (C.c_void_p * 12)(*[int(p) for p in ptrs])
            self._check(self._L.noahmp_b200_bind_forcing(self._ctx, arr))

    def set_fetch(self, fields=()):
        """Fields every RESIDENT-mode noahmplsm() call refreshes on the host (pipelined with the step)."""
        self._check(self._L.noahmp_b200_set_fetch(self._ctx, ",".join(fields).encode()))

    def set_push(self, fields=()):
        """INOUT fields whose host content every RESIDENT-mode noahmplsm() call takes again before the step (the HRLDAS
        driver rewrites LAI = XLAIXY from the forcing file before every call)."""
        self._check_rc(self._L.noahmp_b200_set_push(self._ctx, ",".join(fields).encode()))

    def set_forcing_hints(self, hints):
        """Bit-or of HINT_DZ8W_CONSTANT, HINT_VEGFRA_UNCHANGED, HINT_P8W_LEVELS_EQUAL: forcing planes the following
        RESIDENT-mode calls need not upload again."""
        self._check_rc(self._L.noahmp_b200_set_forcing_hints(self._ctx, int(hints)))

    def set_rebin(self, interval):
        """Re-bin land columns every `interval` RESIDENT-mode steps (0 = never)."""
        self._check(self._L.noahmp_b200_set_rebin(self._ctx, interval))

    @property
    def rebins(self):
        return self._L.noahmp_b200_rebin_count(self._ctx)

    def set_chunks(self, n):
        self._check(self._L.noahmp_b200_set_chunks(self._ctx, n))

    def fetch(self, arrays, scalars, field):
        a = _capi.make_args(self._hold(arrays), scalars)
        self._check(self._L.noahmp_b200_fetch(self._ctx, C.byref(a), field.encode()))

    def step_device(self, itimestep, yr, julian, dt, stream=None):
        self._check(self._L.noahmp_b200_step_device(self._ctx, itimestep, yr, julian, dt, stream))

    def status(self):
        st = _capi.NoahmpStatus()
        self._check(self._L.noahmp_b200_get_status(self._ctx, C.byref(st)))
        return st

    def device_state(self, field, layer=0):
        n = C.c_longlong()
        p = self._L.noahmp_b200_device_state(self._ctx, field.encode(), layer, C.byref(n))
        if not p:
            raise KeyError(field)
        return _DevArray(p, (n.value,), "<i4" if field in _capi.INT_ARRAYS else "<f4")

    # ---- forcing pipeline on the device (row f2) ---------------------------------------------------------
    def forcing_static(self, lat2d, lon2d, zlvl=30.0):
        self._keep = (np.ascontiguousarray(lat2d, np.float32), np.ascontiguousarray(lon2d, np.float32))
        self._check_rc(self._L.noahmp_b200_forcing_static(self._ctx, self._keep[0].ctypes.data, self._keep[1].ctypes.data,
                                                          zlvl))

    def forcing_upload(self, slot, fields):
        """fields: dict t q u v p lw sw pcp fpar of (nj, ni) float32 arrays = one forcing file; slot 0 = A, 1 = B.
        The copy is asynchronous: keep the arrays alive and unchanged until the next forcing_apply returns."""
        f = _capi.make_forcing_fields(self._hold(fields))
        self._check_rc(self._L.noahmp_b200_forcing_upload(self._ctx, slot, C.byref(f)))

    def forcing_swap(self):
        self._check_rc(self._L.noahmp_b200_forcing_swap(self._ctx))

    def forcing_apply(self, fraction, iday, ihour, iminute, isecond, dt):
        j = C.c_float()
        self._check_rc(self._L.noahmp_b200_forcing_apply(self._ctx, fraction, iday, ihour, iminute, isecond, dt,
                                                         C.byref(j)))
        return j.value

    def noahmplsm_device_forcing(self, arrays, scalars):
        a = _capi.make_args(self._hold(arrays), scalars)
        st = _capi.NoahmpStatus()
        self._check(self._L.noahmp_b200_noahmplsm_device_forcing(self._ctx, C.byref(a), C.byref(st)))
        return st

    # ---- output / restart staging -----------------------------------------------------------------------
    def output_begin(self, arrays, scalars, fields="*", mask_water=True):
        """Snapshot `fields` (list or comma separated names, "*" = all state arrays) as of the latest step and start
        copying them into `arrays` in the background; water points become -1.E33 when mask_water (history output)."""
        self._out_args = _capi.make_args(self._hold(arrays), scalars)  # keep the struct and the arrays alive until output_wait
        self._out_keep = arrays
        f = fields if isinstance(fields, str) else ",".join(fields)
        self._check_rc(self._L.noahmp_b200_output_begin(self._ctx, C.byref(self._out_args), f.encode(), int(bool(mask_water))))

    def output_wait(self):
        self._check_rc(self._L.noahmp_b200_output_wait(self._ctx))
        self._out_args = self._out_keep = None

    # ---- cold start (NOAHMP_INIT) ----------------------------------------------------------------------
    def init(self, arrays, scalars):
        """CALL NOAHMP_INIT(...): fills the state arrays in place for a cold start (restart=0); the arrays named in
        _capi.INIT_GW are needed only with iopt_run=5.  Returns STEPWTD (iopt_run=5) or None."""
        arrays = dict(arrays)
        step = np.zeros(1, np.int32)
        if scalars.get("iopt_run") == 5:
            arrays["stepwtd"] = step
        a = _capi.make_init_args(arrays, scalars)
        self._check_rc(self._L.noahmp_b200_init(self._ctx, C.byref(a)))
        return int(step[0]) if scalars.get("iopt_run") == 5 else None

    # ---- opt_run = 5 groundwater (WTABLE_mmf_noahmp) --------------------------------------------------
    def wtable(self, arrays, scalars):
        """CALL WTABLE_mmf_noahmp(...) on a tile that needs no halo (single tile = whole domain)."""
        a = _capi.make_wtable_args(self._hold(arrays), scalars)
        self._check_rc(self._L.noahmp_b200_wtable(self._ctx, C.byref(a)))

    def wtable_device(self, arrays, scalars, stream=None):
        """WTABLE_mmf_noahmp enqueued on `stream` (a cudaStream_t handle) behind the preceding step_device(): pass 1,
        the NCCL halo exchange when a communicator is set, pass 2 and the column update; the host does not wait.
        The marshalled argument list is cached per (arrays, scalars) pair."""
        key = (id(arrays), id(scalars))
        if getattr(self, "_wt_cache", (None,))[0] != key:
            self._wt_cache = (key, _capi.make_wtable_args(self._hold(arrays), scalars), arrays)
        self._check_rc(self._L.noahmp_b200_wtable_device(self._ctx, C.byref(self._wt_cache[1]), stream))

    def wtable_exchange(self, stream=None):
        self._check_rc(self._L.noahmp_b200_wtable_exchange(self._ctx, stream))

    # ---- communicator of the tiles of one domain (NCCL inside the library) -----------------------------------
    def comm_unique_id(self):
        """128-byte NCCL id (numpy uint8) made by rank 0; the host program carries it to the other ranks."""
        buf = np.zeros(128, np.uint8)
        self._check_rc(self._L.noahmp_b200_comm_unique_id(buf.ctypes.data))
        return buf

    def comm_init(self, uid, rank, nranks):
        uid = np.ascontiguousarray(uid, np.uint8)
        assert uid.size == 128
        self._check_rc(self._L.noahmp_b200_comm_init(self._ctx, uid.ctypes.data, int(rank), int(nranks)))

    def comm_neighbours(self):
        nb = (C.c_int * 4)()
        self._check_rc(self._L.noahmp_b200_comm_neighbours(self._ctx, nb))
        return tuple(nb)

    # ---- global water / energy budget -------------------------------------------------------------------------
    BUDGET_NAMES = ("storage_mm", "precip_mm", "et_mm", "runoff_mm", "erreng_wm2", "swe_mm", "columns", "steps")

    def budget_enable(self, on=True):
        self._check_rc(self._L.noahmp_b200_budget_enable(self._ctx, int(bool(on))))

    def budget_read(self, global_sum=False, reset=False):
        """dict of the eight fp64 sums (see include/noahmp_b200.h); global_sum: all-reduced over the communicator."""
        out = (C.c_double * 8)()
        self._check_rc(self._L.noahmp_b200_budget_read(self._ctx, out, int(bool(global_sum)), int(bool(reset))))
        return dict(zip(self.BUDGET_NAMES, list(out)))

    def wtable_begin(self, arrays, scalars):
        a = _capi.make_wtable_args(self._hold(arrays), scalars)
        self._check_rc(self._L.noahmp_b200_wtable_begin(self._ctx, C.byref(a)))

    def wtable_end(self, arrays, scalars):
        a = _capi.make_wtable_args(self._hold(arrays), scalars)
        self._check_rc(self._L.noahmp_b200_wtable_end(self._ctx, C.byref(a)))

    def wtable_halo(self):
        """(KCELL, HEAD) device planes of shape (nj+2, ni+2) with the one-cell halo ring."""
        k, h = C.c_void_p(), C.c_void_p()
        self._check_rc(self._L.noahmp_b200_wtable_halo(self._ctx, C.byref(k), C.byref(h)))
        return _DevArray(k.value, (self.nj + 2, self.ni + 2)), _DevArray(h.value, (self.nj + 2, self.ni + 2))

    def wtable_sync_host(self, arrays, scalars):
        a = _capi.make_wtable_args(self._hold(arrays), scalars)
        self._check_rc(self._L.noahmp_b200_wtable_sync_host(self._ctx, C.byref(a)))

    def _check_rc(self, rc):
        if rc:
            raise NoahmpError(rc, self._L.noahmp_b200_last_error().decode())

    # ---- bookkeeping -------------------------------------------------------------------------------
    @property
    def launches(self):
        return self._L.noahmp_b200_launch_count(self._ctx)

    @property
    def variant(self):
        return self._L.noahmp_b200_kernel_variant(self._ctx).decode()

    def census(self):
        c = (C.c_int64 * 4)()
        self._check(self._L.noahmp_b200_census(self._ctx, c))
        return dict(land=c[0], glacier=c[1], seaice=c[2], water=c[3])

    def column_map(self):
        n = sum(v for k, v in self.census().items() if k != "water")
        out = np.empty(n, np.int32)
        self._check(self._L.noahmp_b200_column_map(self._ctx, out.ctypes.data_as(C.POINTER(C.c_int32)), n))
        return out

    def enable_iteration_counts(self, on=True):
        self._check(self._L.noahmp_b200_enable_iteration_counts(self._ctx, int(on)))

    def iteration_counts(self):
        out = np.empty((self.nj, self.ni), np.int32)
        self._check(self._L.noahmp_b200_get_iteration_counts(self._ctx, out.ctypes.data_as(C.POINTER(C.c_int32))))
        return out


class NoahMPDomain:
    """One host process, several GPUs: the process holds the WHOLE-domain arrays, tile r of the
    mpp_land partition lives on devices[r], and every call slices the tiles straight out of / into the global arrays
    (noahmp_b200_domain_*).  `devices` may name a GPU more than once (several tiles per GPU)."""

    def __init__(self, tables, ni, nj, devices, sync=SYNC_FULL, math=MATH_FAST, fetch=None, push=None, hints=0):
        self._L = _lib.lib()
        if isinstance(tables, dict):
            tables = _capi.tables_from_dict(tables)
        self.tables, self.ni, self.nj, self.devices = tables, ni, nj, list(devices)
        dev = (C.c_int * len(self.devices))(*self.devices)
        self._d = self._L.noahmp_b200_domain_create(C.byref(tables), ni, nj, len(self.devices), dev)
        if not self._d:
            raise NoahmpError(100, self._L.noahmp_b200_last_error().decode())
        rc = self._L.noahmp_b200_domain_configure(
            self._d, sync, math, None if fetch is None else ",".join(fetch).encode(),
            None if push is None else ",".join(push).encode(), int(hints))
        if rc:
            raise NoahmpError(rc, self._L.noahmp_b200_last_error().decode())
        self._keep = None

    @property
    def ntiles(self):
        return self._L.noahmp_b200_domain_ntiles(self._d)

    def tile_bounds(self, r):
        v = [C.c_int() for _ in range(4)]
        self._L.noahmp_b200_domain_tile_bounds(self._d, r, *[C.byref(x) for x in v])
        return tuple(x.value for x in v)

    def _check(self, rc):
        if rc in (100, 101, 8):
            raise NoahmpError(rc, self._L.noahmp_b200_last_error().decode())
        return rc

    def noahmplsm(self, arrays, scalars):
        """CALL noahmplsm(...) with the global arrays; scalars carry the domain as memory bounds (ims..jme)."""
        self._keep = arrays
        a = _capi.make_args(arrays, scalars)
        st = _capi.NoahmpStatus()
        self._check(self._L.noahmp_b200_domain_noahmplsm(self._d, C.byref(a), C.byref(st)))
        return st

    def sync_host(self, arrays, scalars):
        a = _capi.make_args(arrays, scalars)
        self._check(self._L.noahmp_b200_domain_sync_host(self._d, C.byref(a)))

    def close(self):
        if getattr(self, "_d", None):
            self._L.noahmp_b200_domain_destroy(self._d)
            self._d = None

    __del__ = close
