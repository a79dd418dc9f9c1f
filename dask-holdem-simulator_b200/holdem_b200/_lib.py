"""ctypes binding of libholdem_b200.so (the C ABI declared in include/holdem_b200.h).

There is no CPU fallback: if the shared library is missing, or no sm_100 device is present, every compute
entry point raises.  Building is explicit (`python -c "import __graft_entry__ as g; g.build()"` or
`make -C dask-holdem-simulator_b200/csrc`); nothing is compiled at import time.
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libholdem_b200.so")

ABI_VERSION = 2
MODE_TREYS = 0
MODE_REFERENCE = 1
HP_ROW = 16
HS_ROW = 3
BAD_ROW = 0xFFFFFFFF

ERR_NO_DEVICE = -1
ERR_BAD_ARG = -2
ERR_CUDA = -3
ERR_NOT_INIT = -4
ERR_BAD_INPUT = -5
ERR_INTERNAL = -6

# every symbol include/holdem_b200.h declares: name -> (restype, argtypes)
_vp, _i, _i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
SYMBOLS = {
    "holdem_abi_version": (_i, []),
    "holdem_last_error": (ctypes.c_char_p, []),
    "holdem_init": (_i, [_i]),
    "holdem_selftest_tables": (_i, []),
    "holdem_sm_count": (_i, [_i]),
    "holdem_eval_batch": (_i, [_vp, _i, _i64, _vp, _vp]),
    "holdem_eval_batch_strict": (_i, [_vp, _i, _i64, _vp, _vp]),
    "holdem_eval_batch_host": (_i, [_vp, _i, _i64, _vp]),
    "holdem_exhaustive7": (_i, [_vp, _vp, _vp]),
    "holdem_exhaustive7_host": (_i, [_vp, _vp]),
    "holdem_category_batch": (_i, [_vp, _i, _i, _i64, _vp, _vp]),
    "holdem_category_batch_host": (_i, [_vp, _i, _i, _i64, _vp]),
    "holdem_hs_batch": (_i, [_vp, _i64, _i, _vp, _vp]),
    "holdem_hs_batch_host": (_i, [_vp, _i64, _i, _vp]),
    "holdem_hp_batch": (_i, [_vp, _i64, _i, _vp, _vp]),
    "holdem_hp_batch_host": (_i, [_vp, _i64, _i, _vp]),
    "holdem_hp_batch_host_staged": (_i, [_vp, _i64, _i, _vp]),
    "holdem_hp_batch_variant": (_i, [_vp, _i64, _i, _i, _i, _vp, _vp]),
    "holdem_hp_ref_variant": (_i, [_vp, _i64, _i, _vp, _vp]),
    "holdem_hs_batch_variant": (_i, [_vp, _i64, _i, _i, _vp, _vp]),
    "holdem_exhaustive7_variant": (_i, [_i, _vp, _vp, _vp]),
    "holdem_rows_hash": (_i, [_vp, _i64, _i, _i64, _vp, _vp]),
    "holdem_decide_batch": (_i, [_vp, _vp, _vp, _vp, _i, _i64, _vp, _vp]),
    "holdem_decide_from_counts": (_i, [_vp, _vp, _i, _i64, _vp, _vp, _vp]),
    "holdem_pack_rank_major": (_i, [_vp, _vp, _i, _i64, _vp, _vp]),
    "holdem_kernel_launches": (ctypes.c_ulonglong, []),
    "holdem_table_showdown": (_i, [_vp, _vp, _vp, _i64, _i, _vp, _vp, _vp]),
    "holdem_table_hp_batch": (_i, [_vp, _vp, _i64, _i, _i, _i, _vp, _vp]),
    "holdem_hp_batch_direct": (_i, [_vp, _i64, _i, _vp, _vp]),
    "holdem_hp_batch_peers": (_i, [_vp, _i64, _i, _vp, ctypes.POINTER(_vp), _i, _vp]),
    "holdem_hp_flop_variant": (_i, [_vp, _i64, _i, _vp, _vp]),
    "holdem_hp_flop_phase_cycles": (_i, [_vp, _i64, _vp, _vp, _vp]),
    "holdem_hp_turn_variant": (_i, [_vp, _i64, _i, _vp, _vp]),
    "holdem_microbench_smem_gather": (_i, [_i, _i, _i, _i, _i, ctypes.POINTER(ctypes.c_float),
                                           ctypes.POINTER(ctypes.c_double), _vp]),
    "holdem_microbench_int_issue": (_i, [_i, _i, _i, _i, ctypes.POINTER(ctypes.c_float),
                                         ctypes.POINTER(ctypes.c_double), _vp]),
}


class HoldemError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"libholdem_b200 error {code}: {message}")
        self.code = code


_lib = None
_lock = threading.Lock()
_inited = set()


def load():
    """Load the shared library (no device needed).  Raises if it has not been built."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise HoldemError(ERR_INTERNAL, f"{LIB_PATH} is missing: build it with __graft_entry__.build() "
                                                "(there is no CPU fallback)")
            lib = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in SYMBOLS.items():
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            if lib.holdem_abi_version() != ABI_VERSION:
                raise HoldemError(ERR_INTERNAL, "ABI version mismatch")
            _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        raise HoldemError(rc, load().holdem_last_error().decode("utf-8", "replace"))


def init(device: int):
    """holdem_init(device) once per device; makes `device` the calling thread's current CUDA device."""
    lib = load()
    check(lib.holdem_init(int(device)))
    _inited.add(int(device))
    return lib
