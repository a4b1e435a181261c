"""ctypes binding of libhopb200.so (include/hopb200.h).

There is no CPU fallback: if the library is missing, or no CUDA device is present when a handle is
requested, this raises.  PyTorch tensors are used by the callers only as device buffers; this
module passes plain pointers.
"""
from __future__ import annotations

import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhopb200.so")

HOPB_OK = 0
HOPB_E_INVALID = -1
HOPB_E_CUDA = -2
HOPB_E_CAPACITY = -3
HOPB_E_TOO_MANY_SITES = -4
HOPB_E_NOMEM = -5
HOPB_MAX_SITES = 32767
SUMMARY_INTS = 11
STATE_INTS = 4


class HopbError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("libhopb200 error %d: %s" % (code, message))
        self.code = code


class TooManySitesError(OverflowError):
    """More than 32767 sites: hop's int16 site map cannot represent them (hop/sitemap.py:518-519)."""


class SynthParamsC(ctypes.Structure):
    _fields_ = [("seed", ctypes.c_uint64), ("n_atoms", ctypes.c_int64), ("n_frames", ctypes.c_int64),
                ("site_stride", ctypes.c_int32), ("n_sites", ctypes.c_int32), ("box", ctypes.c_float),
                ("cube_half", ctypes.c_float), ("sigma_site", ctypes.c_float), ("sigma_bulk", ctypes.c_float),
                ("p_hop", ctypes.c_float), ("p_rebind", ctypes.c_float)]


_vp = ctypes.c_void_p
_i64 = ctypes.c_int64
_u64 = ctypes.c_uint64
_i32 = ctypes.c_int32
_int = ctypes.c_int
_dbl = ctypes.c_double

# name -> argtypes (restype is int unless listed in _RESTYPES)
_SIGNATURES = {
    "hopb_version": [],
    "hopb_launch_count": [],
    "hopb_create": [_int, ctypes.POINTER(_vp)],
    "hopb_destroy": [_vp],
    "hopb_last_error": [_vp],
    "hopb_set_hist_l2_budget": [_vp, _i64],
    "hopb_dev_alloc": [_vp, _u64, ctypes.POINTER(_vp)],
    "hopb_dev_free": [_vp, _vp],
    "hopb_memset": [_vp, _vp, _vp, _int, _u64],
    "hopb_h2d": [_vp, _vp, _vp, _vp, _u64],
    "hopb_d2h": [_vp, _vp, _vp, _vp, _u64],
    "hopb_stream_sync": [_vp, _vp],
    "hopb_grid_create": [_vp, _vp, _int, _vp, _int, _vp, _int, _vp, ctypes.POINTER(_vp)],
    "hopb_grid_destroy": [_vp],
    "hopb_hist_accumulate": [_vp, _vp, _vp, _vp, _i64, _vp],
    "hopb_hist_bulk": [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _dbl, _vp, _vp, _vp, _vp],
    "hopb_density_from_counts": [_vp, _vp, _vp, _vp, _i64, _dbl, _vp],
    "hopb_density_from_counts_ex": [_vp, _vp, _vp, _vp, _i64, _dbl, _vp, _int],
    "hopb_label_sites": [_vp, _vp, _int, _int, _int, _vp, _dbl, _int, _vp, ctypes.POINTER(_i32), _vp],
    "hopb_label_sites_async": [_vp, _vp, _int, _int, _int, _vp, _dbl, _int, _vp, _vp, _vp],
    "hopb_label_sites_counts_async": [_vp, _vp, _vp, _vp, _i64, ctypes.c_double, ctypes.c_double, _int, _vp, _vp, _vp],
    "hopb_grid_width_classes": [_vp],
    "hopb_grid_set_periodic_box": [_vp, _vp, _vp],
    "hopb_insert_bulk": [_vp, _vp, _i64, _vp, _vp, _i32, _vp, _vp],
    "hopb_labels_to_map16": [_vp, _vp, _i64, _vp, _vp],
    "hopb_site_properties": [_vp, _vp, _vp, _vp, _vp, _vp, _int, _vp, _vp, _vp, _vp],
    "hopb_site_map_fused": [_vp, _vp, _vp, _vp, _i64, _dbl, _dbl, _dbl, _int, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                            _vp, _vp, _int, _vp, _vp, _vp, _vp],
    "hopb_rate_fit": [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _dbl, ctypes.c_uint, _vp, _vp],
    "hopb_site_occupancy": [_vp, _vp, _vp, _i64, _i64, _int, _vp],
    "hopb_site_distance_hist": [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _vp, _vp, _int, _vp, _int, _int, _vp],
    "hopb_site_cells": [_vp, _vp, _i64, _vp, _vp, _int, _vp, _i64, _vp, ctypes.POINTER(_i64)],
    "hopb_hist_accumulate_cells": [_vp, _vp, _vp, _vp, _i64, _vp, _vp],
    "hopb_hoptraj": [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _int, _vp, _vp, _vp, _int, _i64, _vp, _vp, _vp, _u64],
    "hopb_hoptraj_strided": [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _int, _vp, _vp, _vp, _int, _i64, _vp, _vp,
                             _vp, _u64],
    "hopb_hop_fixup_strided": [_vp, _vp, _i64, _i64, _int, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _u64],
    "hopb_hoptraj_i16": [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _int, _vp, _vp, _vp, _int, _i64, _vp, _vp,
                         _vp, _u64],
    "hopb_hop_fixup_i16": [_vp, _vp, _i64, _i64, _int, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _u64],
    "hopb_all_transitions_i16": [_vp, _vp, _vp, _i64, _i64, _int, _vp],
    "hopb_hist_accumulate_multi": [_vp, _vp, _int, _vp, _vp, _vp, _i64, _i64, _vp, _vp],
    "hopb_brick_map": [_vp, _vp, _int, _int, _int, _vp, _vp, _vp, _vp, _vp],
    "hopb_hopscan_labels": [_vp, _vp, _vp, _i64, _i64, _i64, _int, _i64, _vp, _vp, _vp, _u64],
    "hopb_hop_combine": [_vp, _vp, _i64, _int, _i64, _i64, _vp, _vp],
    "hopb_hop_advance": [_vp, _vp, _i64, _int, _vp, _int, _vp],
    "hopb_hop_fixup": [_vp, _vp, _i64, _int, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _u64],
    "hopb_events_finalize": [_vp, _vp, _vp, _u64, _i64, _i64, _int, _vp, _vp, _vp, _vp, _vp, ctypes.POINTER(_i64)],
    "hopb_events_finalize_ex": [_vp, _vp, _vp, _u64, _i64, _i64, _int, _vp, _vp, _vp, _vp, _vp, _vp, ctypes.POINTER(_i64)],
    "hopb_events_finalize_dev": [_vp, _vp, _vp, _vp, _u64, _i64, _i64, _vp, _int, _vp, _vp, _vp, _vp, _vp, _vp, _int, _vp],
    "hopb_labels_to_int16": [_vp, _vp, _vp, _i64, _vp],
    "hopb_nvls_allreduce_u32": [_vp, _vp, _vp, _i64, _int, _int],
    "hopb_p2p_allreduce_u32": [_vp, _vp, _vp, _vp, _i64, _int, _int],
    "hopb_nvls_allgather": [_vp, _vp, _vp, _vp, _i64, _int, _int],
    "hopb_all_transitions": [_vp, _vp, _vp, _i64, _i64, _int, _vp],
    "hopb_synth_generate": [_vp, _vp, ctypes.POINTER(SynthParamsC), _vp, _i64, _i64, _vp],
    "hopb_synth_generate_chunk": [_vp, _vp, ctypes.POINTER(SynthParamsC), _vp, _i64, _i64, _vp, _vp, _int],
}
_RESTYPES = {"hopb_launch_count": ctypes.c_uint64, "hopb_destroy": None, "hopb_grid_destroy": None, "hopb_last_error": ctypes.c_char_p}

_lib = None
_lock = threading.RLock()
_handles = {}


def exported_symbols():
    """Names every entry point include/hopb200.h declares (used by the symbol test)."""
    return sorted(_SIGNATURES)


def lib():
    """The loaded shared library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise ImportError(
                        "hop_b200: %s is missing -- build it with `python -m hop_b200.csrc.build` "
                        "(there is no CPU fallback)" % LIB_PATH)
                l = ctypes.CDLL(LIB_PATH)
                for name, argtypes in _SIGNATURES.items():
                    fn = getattr(l, name)
                    fn.argtypes = argtypes
                    fn.restype = _RESTYPES.get(name, ctypes.c_int)
                _lib = l
    return _lib


_lane_base = 0


class lanes(object):
    """Context manager: every handle requested inside uses lane `base` + its own lane number, i.e. a
    set of handles (scratch memory) of its own.  Two pipelines whose work is in flight at the same time
    on different streams each run under their own base."""

    def __init__(self, base):
        self.base = int(base)

    def __enter__(self):
        global _lane_base
        self.saved, _lane_base = _lane_base, self.base
        return self

    def __exit__(self, *exc):
        global _lane_base
        _lane_base = self.saved
        return False


def handle(device_index=None, lane=0):
    """One hopb_handle per CUDA device (created on first use).  `lane` > 0 gives an additional
    handle with its own scratch memory, for work enqueued concurrently on a second stream."""
    import torch

    lane = int(lane) + _lane_base

    if not torch.cuda.is_available():
        raise RuntimeError("hop_b200 needs a CUDA device (B200, sm_100a); no CPU fallback exists")
    if device_index is None:
        device_index = torch.cuda.current_device()
    key = (os.getpid(), int(device_index), int(lane))
    library = lib()
    h = _handles.get(key)
    if h is None:
        with _lock:
            h = _handles.get(key)
            if h is None:
                torch.cuda.init()
                out = _vp()
                rc = library.hopb_create(int(device_index), ctypes.byref(out))
                if rc != 0:
                    raise HopbError(rc, "hopb_create failed on device %d" % device_index)
                h = out
                _handles[key] = h
    return h


def check(h, rc):
    if rc == HOPB_OK:
        return
    msg = lib().hopb_last_error(h)
    msg = msg.decode("utf-8", "replace") if msg else ""
    if rc == HOPB_E_TOO_MANY_SITES:
        raise TooManySitesError(msg)
    if rc == HOPB_E_INVALID:
        raise ValueError(msg)
    raise HopbError(rc, msg)
