# -*- coding: utf-8 -*-
"""ctypes binding of ``libcmpy_b200.so`` (C ABI declared in ``include/cmpy_b200.h``).

There is deliberately no CPU fallback: if the shared library is missing, or no CUDA
device is visible when a compute entry point is called, a ``RuntimeError`` is raised.
PyTorch is used only for device memory, streams and ``torch.distributed`` plumbing.
"""
import ctypes
import os
from ctypes import (POINTER, c_char_p, c_double, c_int, c_int32, c_int64, c_void_p)

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcmpy_b200.so")

_lib = None

CMPY_OK = 0
CMPY_ERR_ARG = -1
CMPY_ERR_CUDA = -2
CMPY_ERR_NOMEM = -3
CMPY_ERR_UNSUPPORTED = -4
CMPY_ERR_NOT_CONVERGED = -5

_p = c_void_p  # device pointers and streams travel as integers

# name -> (restype, argtypes); every symbol declared in include/cmpy_b200.h
SIGNATURES = {
    "cmpy_last_error": (c_char_p, []),
    "cmpy_version": (c_int, []),
    "cmpy_launch_count": (c_int64, []),
    "cmpy_reset_launch_count": (None, []),
    "cmpy_device_info": (c_int, [POINTER(c_int), POINTER(c_int64), POINTER(c_int64)]),
    "cmpy_binomial": (c_int, [c_int, c_int, POINTER(c_int64)]),
    "cmpy_sector_enumerate": (c_int, [c_int, c_int, _p, _p]),
    "cmpy_sector_rank": (c_int, [_p, c_int64, _p, _p]),
    "cmpy_species_hops": (c_int, [_p, c_int64, c_int, c_int, c_int, c_int, _p, _p, _p]),
    "cmpy_weighted_elements": (c_int, [_p, c_int64, POINTER(c_double), c_int, _p, _p]),
    "cmpy_inter_elements": (c_int, [_p, c_int64, _p, c_int64, POINTER(c_double), c_int, _p, _p]),
    "cmpy_hubbard_create": (c_int, [c_int, POINTER(c_int64), c_int64, POINTER(c_int64), c_int64,
                                    c_int, c_int, POINTER(c_int32), POINTER(c_double),
                                    POINTER(c_double), POINTER(c_double), c_int, POINTER(_p)]),
    "cmpy_heisenberg_create": (c_int, [c_int, c_int, c_int, POINTER(c_int32), c_double, c_double,
                                       POINTER(_p)]),
    "cmpy_coo_create": (c_int, [c_int64, c_int64, POINTER(c_int64), POINTER(c_int64),
                                POINTER(c_double), POINTER(_p)]),
    "cmpy_op_destroy": (c_int, [_p]),
    "cmpy_op_size": (c_int, [_p, POINTER(c_int64)]),
    "cmpy_hv_apply": (c_int, [_p, _p, _p, _p]),
    "cmpy_hubbard_apply_rows": (c_int, [_p, _p, _p, c_int64, c_int64, c_int, _p]),
    "cmpy_hv_set_variant": (c_int, [_p, c_int]),
    "cmpy_op_trace": (c_int, [_p, POINTER(c_double)]),
    "cmpy_op_diagonal": (c_int, [_p, _p, _p]),
    "cmpy_ladder_apply": (c_int, [_p, c_int64, _p, c_int64, _p, c_int64, _p, c_int64, c_int, c_int,
                                  c_int, c_int, c_int, _p, _p, _p]),
    "cmpy_dot": (c_int, [_p, _p, _p, c_int64, _p, _p]),
    "cmpy_lanczos_run": (c_int, [_p, _p, _p, _p, c_int, c_double, c_double, c_int, c_int,
                                 POINTER(c_double), POINTER(c_double), POINTER(c_int),
                                 POINTER(c_double), POINTER(c_double), _p, _p]),
    "cmpy_tridiag_lowest": (c_int, [POINTER(c_double), POINTER(c_double), c_int, c_int,
                                    POINTER(c_double), POINTER(c_double)]),
    "cmpy_cf_eval": (c_int, [POINTER(c_double), POINTER(c_double), c_int, c_double, c_double, c_int,
                             _p, c_int64, _p, c_int, _p]),
    "cmpy_pole_sum": (c_int, [_p, _p, c_int64, _p, c_int64, _p, c_int, _p]),
    "cmpy_transpose": (c_int, [_p, c_int64, c_int64, c_int64, _p, c_int64, c_int, _p]),
    "cmpy_copy2d": (c_int, [_p, c_int64, c_int64, c_int64, _p, c_int64, c_int, _p]),
    "cmpy_hubbard_set_grid_limit": (c_int, [_p, c_int]),
    "cmpy_transpose_push": (c_int, [_p, c_int64, c_int64, c_int64, c_int64, c_int, POINTER(c_int64),
                                    POINTER(_p), _p]),
    "cmpy_transpose_push_capped": (c_int, [_p, c_int64, c_int64, c_int64, c_int64, c_int, POINTER(c_int64),
                                           POINTER(_p), c_int, _p]),
    "cmpy_transpose_pull_acc": (c_int, [_p, c_int64, c_int64, c_int64, c_int64, c_int, POINTER(c_int64),
                                        POINTER(_p), _p]),
    "cmpy_dist_ctl_bytes": (c_int, []),
    "cmpy_dist_create": (c_int, [_p, _p, c_int, c_int, POINTER(_p), POINTER(_p), POINTER(_p), POINTER(_p)]),
    "cmpy_dist_destroy": (c_int, [_p]),
    "cmpy_hv_apply_sharded": (c_int, [_p, _p, _p, c_int, _p]),
    "cmpy_dist_allreduce_sum": (c_int, [_p, _p, _p, _p]),
    "cmpy_dist_barrier": (c_int, [_p, _p]),
    "cmpy_lanczos_sharded": (c_int, [_p, _p, _p, c_int, c_double, c_int, POINTER(c_double), POINTER(c_double),
                                     POINTER(c_int), POINTER(c_double), _p]),
}


class CmpyError(RuntimeError):
    """Error reported by libcmpy_b200 (CUDA failure, unsupported configuration...)."""


class NotConverged(CmpyError):
    pass


def lib():
    """Loads (once) and returns the ctypes handle of libcmpy_b200.so."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; "
                f"g.build()'` or `make -C cmpy_b200/csrc`. cmpy_b200 has no CPU fallback.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def last_error():
    msg = lib().cmpy_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc, what=""):
    """Maps a C status to the exception type the reference would raise."""
    if rc == CMPY_OK:
        return
    msg = f"{what}: {last_error()}" if what else last_error()
    if rc == CMPY_ERR_ARG:
        raise ValueError(msg)
    if rc == CMPY_ERR_NOMEM:
        raise MemoryError(msg)
    if rc == CMPY_ERR_NOT_CONVERGED:
        raise NotConverged(msg)
    raise CmpyError(msg)


def torch_mod():
    import torch

    return torch


def require_cuda():
    torch = torch_mod()
    if not torch.cuda.is_available():
        raise RuntimeError("cmpy_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    lib()
    return torch


def device():
    torch = require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


def stream_ptr():
    torch = torch_mod()
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device pointer of a torch CUDA tensor (or None)."""
    if t is None:
        return c_void_p(0)
    return c_void_p(t.data_ptr())


def as_c_array(arr, ctype, dtype):
    a = np.ascontiguousarray(arr, dtype=dtype)
    return a, a.ctypes.data_as(POINTER(ctype))


def launch_count():
    return int(lib().cmpy_launch_count())


def reset_launch_count():
    lib().cmpy_reset_launch_count()


def binomial(n, k):
    out = c_int64(0)
    check(lib().cmpy_binomial(int(n), int(k), ctypes.byref(out)), "cmpy_binomial")
    return int(out.value)
