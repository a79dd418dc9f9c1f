"""Batched host-side API over the C ABI (include/holdem_b200.h).

Device entry points take CUDA tensors, are ordered on torch's current stream and never synchronise; host entry
points take numpy arrays and go through the library's own pinned-staging pipeline.  PyTorch is used for device
memory and streams only -- every number comes from the hand-written kernels.  No CPU fallback exists: without a
CUDA device (or without the built library) these functions raise.

Integer outputs are returned as signed torch dtypes of the same width (uint16 ranks -> int16, uint32 counts -> int32);
all legitimate values are far below the sign bit and HOLDEM_BAD_ROW reads as -1.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import MODE_REFERENCE, MODE_TREYS, HP_ROW, HS_ROW

__all__ = [
    "MODE_TREYS", "MODE_REFERENCE", "mode_id", "evaluate", "category", "hand_strength_counts",
    "hand_potential_counts", "flop_phase_cycles", "exhaustive7", "hs_from_counts", "potentials_from_counts",
    "normalised_potentials", "ehs_from_counts", "evaluate_host", "category_host", "hand_strength_counts_host",
    "hand_potential_counts_host", "smem_gather_microbench", "int_issue_microbench", "EVALS_PER_FLOP_SITUATION",
]

# algorithmic evaluations per situation
EVALS_PER_FLOP_SITUATION = {MODE_TREYS: 1_072_353, MODE_REFERENCE: 1_273_514}
EVALS_PER_TURN_SITUATION = {MODE_TREYS: 46_622}


def mode_id(mode) -> int:
    if mode in (MODE_TREYS, "treys"):
        return MODE_TREYS
    if mode in (MODE_REFERENCE, "reference"):
        return MODE_REFERENCE
    raise ValueError(f"mode must be 'treys' or 'reference', got {mode!r}")


def _require_cuda(t: torch.Tensor, name: str, dtype, cols=None):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (this path has no CPU fallback)")
    if t.dtype != dtype or not t.is_contiguous():
        raise ValueError(f"{name} must be a contiguous {dtype} tensor")
    if cols is not None and (t.dim() != 2 or t.shape[1] != cols):
        raise ValueError(f"{name} must have shape [N,{cols}]")
    return t


def _enter(t: torch.Tensor):
    """Initialise the library on the tensor's device; returns (lib, raw stream handle)."""
    dev = t.device.index if t.device.index is not None else torch.cuda.current_device()
    torch.cuda.set_device(dev)
    lib = _lib.init(dev)
    return lib, ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


# ---- device entry points ---------------------------------------------------------------------------------------------
def evaluate(cards: torch.Tensor, n_cards: int, out: torch.Tensor = None, strict: bool = False) -> torch.Tensor:
    """Treys rank (1..7462, lower is better) of each row's first n_cards (5/6/7) cards.  cards: uint8 [N,8].
    Rows with a card > 51 give 0.  strict=True also gives 0 for rows with duplicate cards (the large-batch fast path
    does not look for duplicates, exactly like Treys)."""
    _require_cuda(cards, "cards", torch.uint8, 8)
    if n_cards not in (5, 6, 7):
        raise KeyError(n_cards)  # treys.Evaluator.hand_size_map
    if out is None:
        out = torch.empty(cards.shape[0], dtype=torch.int16, device=cards.device)
    lib, st = _enter(cards)
    fn = lib.holdem_eval_batch_strict if strict else lib.holdem_eval_batch
    _lib.check(fn(cards.data_ptr(), n_cards, cards.shape[0], out.data_ptr(), st))
    return out


def category(cards: torch.Tensor, n_cards: int) -> torch.Tensor:
    """Literal CalculateHandRankValue category 0..8 of each row's first n_cards (2..9) cards, duplicates allowed."""
    _require_cuda(cards, "cards", torch.uint8)
    if cards.dim() != 2 or cards.shape[1] < n_cards:
        raise ValueError("cards must be [N, >= n_cards]")
    out = torch.empty(cards.shape[0], dtype=torch.uint8, device=cards.device)
    lib, st = _enter(cards)
    _lib.check(lib.holdem_category_batch(cards.data_ptr(), n_cards, cards.shape[1], cards.shape[0],
                                         out.data_ptr(), st))
    return out


def hand_strength_counts(sit: torch.Tensor, mode="treys") -> torch.Tensor:
    """(ahead, tied, behind) per situation: int32 [N,3].  sit: uint8 [N,8] rows h0 h1 b0..b4(0xFF) pad."""
    _require_cuda(sit, "sit", torch.uint8, 8)
    out = torch.empty((sit.shape[0], HS_ROW), dtype=torch.int32, device=sit.device)
    lib, st = _enter(sit)
    _lib.check(lib.holdem_hs_batch(sit.data_ptr(), sit.shape[0], mode_id(mode), out.data_ptr(), st))
    return out


def hand_potential_counts(sit: torch.Tensor, mode="treys", out: torch.Tensor = None,
                          direct: bool = False, flop_variant: int = None, turn_variant: int = None) -> torch.Tensor:
    """HS + HP counts per situation: int32 [N,16] = ahead tied behind | HP[3][3] | HPTotal[3] | n_board.
    `out` may be a [N,16] int32 view into a larger buffer (e.g. this rank's slice of an all-gather buffer).
    direct=True forces the plain-enumeration kernel (the independent second implementation).  flop_variant (treys mode,
    flop rows only; 4 = rank-multiset kernel, 3 / 2 = the 4-set kernels) runs one named implementation of the flop kernel
    for A/B measurements; rows of other streets are then left as they are in `out`.  turn_variant (4 = rank-multiset
    kernel, 3 = 3-set walk) does the same for turn rows."""
    _require_cuda(sit, "sit", torch.uint8, 8)
    if out is None:
        out = torch.empty((sit.shape[0], HP_ROW), dtype=torch.int32, device=sit.device)
    else:
        _require_cuda(out, "out", torch.int32, HP_ROW)
        if out.shape[0] != sit.shape[0] or out.device != sit.device:
            raise ValueError("out must be [N,16] on the same device as sit")
    lib, st = _enter(sit)
    if turn_variant is not None:
        if mode_id(mode) != MODE_TREYS or direct or flop_variant is not None:
            raise ValueError("turn_variant applies to mode='treys' alone")
        _lib.check(lib.holdem_hp_turn_variant(sit.data_ptr(), sit.shape[0], int(turn_variant), out.data_ptr(), st))
        return out
    if flop_variant is not None:
        if mode_id(mode) != MODE_TREYS or direct:
            raise ValueError("flop_variant applies to mode='treys' without direct=True")
        _lib.check(lib.holdem_hp_flop_variant(sit.data_ptr(), sit.shape[0], int(flop_variant), out.data_ptr(), st))
        return out
    fn = lib.holdem_hp_batch_direct if direct else lib.holdem_hp_batch
    _lib.check(fn(sit.data_ptr(), sit.shape[0], mode_id(mode), out.data_ptr(), st))
    return out


def flop_phase_cycles(sit: torch.Tensor):
    """Production flop kernel with phase clocks on -> (rows int32 [N,16], cycles int64 [4] = tables, W+T, descriptors,
    tasks; SM cycles summed over all situations)."""
    _require_cuda(sit, "sit", torch.uint8, 8)
    out = torch.empty((sit.shape[0], HP_ROW), dtype=torch.int32, device=sit.device)
    cyc = torch.zeros(4, dtype=torch.int64, device=sit.device)
    lib, st = _enter(sit)
    _lib.check(lib.holdem_hp_flop_phase_cycles(sit.data_ptr(), sit.shape[0], out.data_ptr(), cyc.data_ptr(), st))
    return out, cyc


def exhaustive7(device=None):
    """All C(52,7) hands evaluated in-kernel -> (hist int64 [7463] indexed by rank, sums int64 [2] = sum r, sum r^2)."""
    dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
    hist = torch.empty(7463, dtype=torch.int64, device=dev)
    sums = torch.empty(2, dtype=torch.int64, device=dev)
    lib, st = _enter(hist)
    _lib.check(lib.holdem_exhaustive7(hist.data_ptr(), sums.data_ptr(), st))
    return hist, sums


# ---- floats from counts (the reference's formulas, same dtype and operation order) -----------------------------------------
def hs_from_counts(counts: torch.Tensor) -> torch.Tensor:
    """HandCalculations.py:175 in float64 (a Python float is a double): (ahead + tied/2) / (ahead+tied+behind)."""
    c = counts[..., 0:3].to(torch.float64)
    return (c[..., 0] + c[..., 1] / 2) / (c[..., 0] + c[..., 1] + c[..., 2])


def potentials_from_counts(rows: torch.Tensor):
    """HandCalculations.py:221-224 on float32 tensors, the reference's operation order -> (Ppot, Npot).
    As in the reference the 3x3 counts are NOT rescaled by the number of runouts."""
    hp = rows[..., 3:12].to(torch.float32).reshape(*rows.shape[:-1], 3, 3)
    tot = rows[..., 12:15].to(torch.float32)
    ahead, tied, behind = 0, 1, 2
    ppot = (hp[..., behind, ahead] + hp[..., behind, tied] / 2 + hp[..., tied, ahead] / 2) / \
           (tot[..., behind] + tot[..., tied])
    npot = (hp[..., ahead, behind] + hp[..., tied, behind] / 2 + hp[..., ahead, tied] / 2) / \
           (tot[..., ahead] + tot[..., tied])
    return ppot, npot


def normalised_potentials(rows: torch.Tensor):
    """PPot/NPot as probabilities: the reference's formula with each HP row divided by the number of runouts per
    opponent (sum of the row / HPTotal of the row).  Not computed anywhere in the reference (parity unpinned);
    these are the values Decision.py's 0.3/0.6/0.8 thresholds are meaningful for.  float64."""
    hp = rows[..., 3:12].to(torch.float64).reshape(*rows.shape[:-1], 3, 3)
    tot = rows[..., 12:15].to(torch.float64)
    runouts = hp.sum(dim=(-1, -2)) / tot.sum(dim=-1)
    runouts = torch.where(runouts > 0, runouts, torch.ones_like(runouts))
    ahead, tied, behind = 0, 1, 2
    ppot = (hp[..., behind, ahead] + hp[..., behind, tied] / 2 + hp[..., tied, ahead] / 2) / \
           ((tot[..., behind] + tot[..., tied]) * runouts)
    npot = (hp[..., ahead, behind] + hp[..., tied, behind] / 2 + hp[..., ahead, tied] / 2) / \
           ((tot[..., ahead] + tot[..., tied]) * runouts)
    return ppot, npot


def ehs_from_counts(rows: torch.Tensor) -> torch.Tensor:
    """Effective hand strength HS*(1-NPot) + (1-HS)*PPot on the normalised potentials (float64; the formula of the
    Wikipedia article HandCalculations.py:14 links to -- the reference never evaluates it, parity unpinned).
    Where a potential is undefined (0/0: nobody ahead/behind) its term contributes with potential 0."""
    hs = hs_from_counts(rows)
    ppot, npot = normalised_potentials(rows)
    ppot = torch.nan_to_num(ppot, nan=0.0, posinf=0.0)
    npot = torch.nan_to_num(npot, nan=0.0, posinf=0.0)
    return hs * (1 - npot) + (1 - hs) * ppot


# ---- host entry points (numpy in / numpy out; copies happen inside the library) -----------------------------------------------
def _host_enter(device):
    dev = torch.cuda.current_device() if device is None else int(device)
    torch.cuda.set_device(dev)
    return _lib.init(dev)


def _np_u8(a, cols=None):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    if a.ndim != 2 or (cols is not None and a.shape[1] != cols):
        raise ValueError(f"expected a uint8 array of shape [N,{cols if cols else 'S'}]")
    return a


def evaluate_host(cards, n_cards: int, device=None) -> np.ndarray:
    cards = _np_u8(cards, 8)
    if n_cards not in (5, 6, 7):
        raise KeyError(n_cards)
    out = np.empty(cards.shape[0], dtype=np.uint16)
    lib = _host_enter(device)
    _lib.check(lib.holdem_eval_batch_host(cards.ctypes.data, n_cards, cards.shape[0], out.ctypes.data))
    return out


def category_host(cards, n_cards: int, device=None) -> np.ndarray:
    cards = _np_u8(cards)
    out = np.empty(cards.shape[0], dtype=np.uint8)
    lib = _host_enter(device)
    _lib.check(lib.holdem_category_batch_host(cards.ctypes.data, n_cards, cards.shape[1], cards.shape[0],
                                              out.ctypes.data))
    return out


def hand_strength_counts_host(sit, mode="treys", device=None) -> np.ndarray:
    sit = _np_u8(sit, 8)
    out = np.empty((sit.shape[0], HS_ROW), dtype=np.uint32)
    lib = _host_enter(device)
    _lib.check(lib.holdem_hs_batch_host(sit.ctypes.data, sit.shape[0], mode_id(mode), out.ctypes.data))
    return out


def hand_potential_counts_host(sit, mode="treys", device=None, out=None) -> np.ndarray:
    sit = _np_u8(sit, 8)
    if out is None:
        out = np.empty((sit.shape[0], HP_ROW), dtype=np.uint32)
    lib = _host_enter(device)
    _lib.check(lib.holdem_hp_batch_host(sit.ctypes.data, sit.shape[0], mode_id(mode), out.ctypes.data))
    return out


def smem_gather_microbench(table_bytes=131072, iters=2048, pattern=0, blocks_per_sm=1, threads=1024, device=None):
    """Measured shared-memory uint16 gather rate (lookups/s) -- the lookup roofline denominator."""
    lib = _host_enter(device)
    ms = ctypes.c_float(0)
    lookups = ctypes.c_double(0)
    _lib.check(lib.holdem_microbench_smem_gather(table_bytes, iters, pattern, blocks_per_sm, threads,
                                                 ctypes.byref(ms), ctypes.byref(lookups), None))
    return lookups.value / (ms.value * 1e-3), ms.value


def int_issue_microbench(mix=0, iters=4096, blocks_per_sm=1, threads=1024, device=None):
    """Measured integer instruction issue rate (warp instructions/s) -- the issue roofline denominator.
    mix 0: compare + predicated multiply-add pairs (ALU + FMA pipes), 1: FMA pipe only, 2: ALU pipe only."""
    lib = _host_enter(device)
    ms = ctypes.c_float(0)
    n = ctypes.c_double(0)
    _lib.check(lib.holdem_microbench_int_issue(iters, mix, blocks_per_sm, threads, ctypes.byref(ms), ctypes.byref(n), None))
    return n.value / (ms.value * 1e-3), ms.value
