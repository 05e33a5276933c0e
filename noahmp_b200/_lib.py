"""ctypes binding of noahmp_b200/libnoahmp_b200.so — the C-ABI declared in include/noahmp_b200.h.

This is the same binding surface the Fortran ISO_C_BINDING shim of INTEGRATION.md uses.  The library is
sm_100a CUDA only: loading fails loudly when it has not been built (python -c 'import __graft_entry__ as g;
g.build()'), and noahmp_b200_create() fails when no CUDA device is present.  There is no CPU fallback.
"""
import ctypes as C
import os
import subprocess

from . import _capi

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("NOAHMP_B200_LIB") or os.path.join(_HERE, "libnoahmp_b200.so")
_LIB = None

_pa = C.POINTER(_capi.NoahmpLsmArgs)
_pt = C.POINTER(_capi.NoahmpTables)
_ps = C.POINTER(_capi.NoahmpStatus)
_pw = C.POINTER(_capi.NoahmpWtableArgs)
_pinit = C.POINTER(_capi.NoahmpInitArgs)
_pff = C.POINTER(_capi.NoahmpForcingFields)
_ctx = C.c_void_p

# name -> (restype, argtypes); every symbol include/noahmp_b200.h declares
SYMBOLS = {
    "noahmp_b200_read_tables": (C.c_int, [C.c_char_p, C.c_char_p, C.c_char_p, _pt]),
    "noahmp_b200_tables_error": (C.c_char_p, []),
    "noahmp_b200_sizeof_tables": (C.c_ulonglong, []),
    "noahmp_b200_sizeof_args": (C.c_ulonglong, []),
    "noahmp_b200_create": (_ctx, [C.c_int, _pt, C.c_int, C.c_int]),
    "noahmp_b200_destroy": (None, [_ctx]),
    "noahmp_b200_last_error": (C.c_char_p, []),
    "noahmp_b200_set_mode": (C.c_int, [_ctx, C.c_int]),
    "noahmp_b200_set_math": (C.c_int, [_ctx, C.c_int]),
    "noahmp_b200_kernel_variant": (C.c_char_p, [_ctx]),
    "noahmp_b200_noahmplsm": (C.c_int, [_ctx, _pa, _ps]),
    "noahmp_b200_sync_host": (C.c_int, [_ctx, _pa]),
    "noahmp_b200_set_fetch": (C.c_int, [_ctx, C.c_char_p]),
    "noahmp_b200_set_push": (C.c_int, [_ctx, C.c_char_p]),
    "noahmp_b200_set_forcing_hints": (C.c_int, [_ctx, C.c_uint]),
    "noahmp_b200_unpin": (C.c_int, [_ctx, C.c_void_p]),
    "noahmp_b200_set_rebin": (C.c_int, [_ctx, C.c_int]),
    "noahmp_b200_rebin_count": (C.c_int, [_ctx]),
    "noahmp_b200_set_chunks": (C.c_int, [_ctx, C.c_int]),
    "noahmp_b200_fetch": (C.c_int, [_ctx, _pa, C.c_char_p]),
    "noahmp_b200_bind_forcing": (C.c_int, [_ctx, C.POINTER(C.c_void_p)]),
    "noahmp_b200_upload": (C.c_int, [_ctx, _pa]),
    "noahmp_b200_device_forcing": (C.c_int, [_ctx, C.POINTER(C.c_void_p)]),
    "noahmp_b200_step_device": (C.c_int, [_ctx, C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p]),
    "noahmp_b200_get_status": (C.c_int, [_ctx, _ps]),
    "noahmp_b200_launch_count": (C.c_longlong, [_ctx]),
    "noahmp_b200_census": (C.c_int, [_ctx, C.POINTER(C.c_int64)]),
    "noahmp_b200_column_map": (C.c_int, [_ctx, C.POINTER(C.c_int32), C.c_longlong]),
    "noahmp_b200_device_state": (C.c_void_p, [_ctx, C.c_char_p, C.c_int, C.POINTER(C.c_longlong)]),
    "noahmp_b200_enable_iteration_counts": (C.c_int, [_ctx, C.c_int]),
    "noahmp_b200_get_iteration_counts": (C.c_int, [_ctx, C.POINTER(C.c_int32)]),
    "noahmp_b200_forcing_static": (C.c_int, [_ctx, C.c_void_p, C.c_void_p, C.c_float]),
    "noahmp_b200_forcing_upload": (C.c_int, [_ctx, C.c_int, _pff]),
    "noahmp_b200_forcing_swap": (C.c_int, [_ctx]),
    "noahmp_b200_forcing_apply": (C.c_int, [_ctx, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                           C.POINTER(C.c_float)]),
    "noahmp_b200_noahmplsm_device_forcing": (C.c_int, [_ctx, _pa, _ps]),
    "noahmp_b200_init": (C.c_int, [_ctx, _pinit]),
    "noahmp_b200_output_begin": (C.c_int, [_ctx, _pa, C.c_char_p, C.c_int]),
    "noahmp_b200_output_wait": (C.c_int, [_ctx]),
    "noahmp_b200_sizeof_init_args": (C.c_ulonglong, []),
    "noahmp_b200_wtable": (C.c_int, [_ctx, _pw]),
    "noahmp_b200_wtable_begin": (C.c_int, [_ctx, _pw]),
    "noahmp_b200_wtable_end": (C.c_int, [_ctx, _pw]),
    "noahmp_b200_wtable_halo": (C.c_int, [_ctx, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "noahmp_b200_wtable_sync_host": (C.c_int, [_ctx, _pw]),
    "noahmp_b200_wtable_exchange": (C.c_int, [_ctx, C.c_void_p]),
    "noahmp_b200_wtable_device": (C.c_int, [_ctx, _pw, C.c_void_p]),
    "noahmp_b200_comm_unique_id": (C.c_int, [C.c_void_p]),
    "noahmp_b200_comm_init": (C.c_int, [_ctx, C.c_void_p, C.c_int, C.c_int]),
    "noahmp_b200_comm_neighbours": (C.c_int, [_ctx, C.POINTER(C.c_int)]),
    "noahmp_b200_tile_neighbours": (None, [C.c_int, C.c_int, C.POINTER(C.c_int)]),
    "noahmp_b200_budget_enable": (C.c_int, [_ctx, C.c_int]),
    "noahmp_b200_budget_read": (C.c_int, [_ctx, C.POINTER(C.c_double), C.c_int, C.c_int]),
    "noahmp_b200_domain_create": (C.c_void_p, [_pt, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]),
    "noahmp_b200_domain_destroy": (None, [C.c_void_p]),
    "noahmp_b200_domain_ntiles": (C.c_int, [C.c_void_p]),
    "noahmp_b200_domain_tile": (C.c_void_p, [C.c_void_p, C.c_int]),
    "noahmp_b200_domain_tile_bounds": (C.c_int, [C.c_void_p, C.c_int] + [C.POINTER(C.c_int)] * 4),
    "noahmp_b200_domain_configure": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_char_p, C.c_char_p, C.c_uint]),
    "noahmp_b200_domain_noahmplsm": (C.c_int, [C.c_void_p, _pa, _ps]),
    "noahmp_b200_domain_sync_host": (C.c_int, [C.c_void_p, _pa]),
    "noahmp_b200_proc_grid": (None, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "noahmp_b200_tile": (None, [C.c_int, C.c_int, C.c_int, C.c_int] + [C.POINTER(C.c_int)] * 4),
}


def build(force=False):
    """Compile the CUDA extension in-tree (nvcc, sm_100a). Cross-compiles without a GPU."""
    src = os.path.join(_HERE, "csrc")
    if not force and os.path.exists(SO_PATH):
        # a prebuilt library that is newer than every source is used as it is, also when the intermediate objects did
        # not travel with it (the GPU box receives the .so, not necessarily csrc/build/)
        deps = [os.path.join(src, f) for f in os.listdir(src) if not os.path.isdir(os.path.join(src, f))]
        deps.append(os.path.join(os.path.dirname(_HERE), "include", "noahmp_b200.h"))
        if os.path.getmtime(SO_PATH) >= max(os.path.getmtime(d) for d in deps):
            return SO_PATH
    args = ["make", "-C", src, "-s", "-j4"]
    if force:
        args.append("-B")
    subprocess.check_call(args)
    return SO_PATH


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError(
                f"{SO_PATH} is missing: build the CUDA extension first (noahmp_b200._lib.build() or "
                "__graft_entry__.build()). noahmp_b200 has no CPU fallback.")
        L = C.CDLL(SO_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError if the library lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB
