"""torch.ops.holdem.* -- the thin TORCH_LIBRARY layer (csrc/torch_ops.cpp) over the C ABI.

`load()` registers the ops from libholdem_b200_torch.so (built in-tree next to libholdem_b200.so); each op takes the
tensor's device guard and torch's current stream itself and allocates its outputs with torch.  Only a CUDA implementation
is registered: CPU tensors raise (no CPU fallback).  `torch.ops.holdem.hp_sharded` is added here in Python because it is
plumbing over torch.distributed (symmetric memory / NCCL), not a kernel call.
"""
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
TORCH_LIB_PATH = os.path.join(_HERE, "libholdem_b200_torch.so")
_loaded = False
_lock = threading.Lock()


def load():
    """Idempotent; raises if the extension has not been built (there is no fallback binding chosen silently)."""
    global _loaded
    with _lock:
        if not _loaded:
            if not os.path.exists(TORCH_LIB_PATH):
                raise RuntimeError(f"{TORCH_LIB_PATH} is missing: build it with __graft_entry__.build()")
            from . import _lib
            _lib.load()                      # the C-ABI library first (same file the extension links against)
            torch.ops.load_library(TORCH_LIB_PATH)
            lib = torch.library.Library("holdem", "FRAGMENT")
            lib.define("hp_sharded(Tensor sit, int mode) -> Tensor")

            def _hp_sharded(sit, mode):
                from .sharded import hand_potential_counts_sharded
                return hand_potential_counts_sharded(sit, int(mode))

            lib.impl("hp_sharded", _hp_sharded, "CUDA")
            load._lib = lib                  # keep the fragment alive
            _loaded = True
    return torch.ops.holdem
