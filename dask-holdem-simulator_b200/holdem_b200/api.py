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

import contextlib

from . import _lib, _ops
from ._lib import MODE_REFERENCE, MODE_TREYS, HP_ROW, HS_ROW

__all__ = [
    "MODE_TREYS", "MODE_REFERENCE", "mode_id", "evaluate", "category", "hand_strength_counts",
    "hand_potential_counts", "flop_phase_cycles", "exhaustive7", "hs_from_counts", "potentials_from_counts",
    "normalised_potentials", "ehs_from_counts", "evaluate_host", "category_host", "hand_strength_counts_host",
    "hand_potential_counts_host", "smem_gather_microbench", "int_issue_microbench", "EVALS_PER_FLOP_SITUATION",
    "rows_hash", "rows_hash_host", "table_hand_potential_counts", "table_showdown", "kernel_launches", "pack_rank_major",
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


@contextlib.contextmanager
def _enter(t: torch.Tensor):
    """ctypes binding: make the tensor's device current for the duration of the C call (the caller's current device is
    restored afterwards), initialise the library there once; yields (lib, raw handle of torch's current stream)."""
    dev = t.device.index if t.device.index is not None else torch.cuda.current_device()
    with torch.cuda.device(dev):
        lib = _lib.load() if dev in _lib._inited else _lib.init(dev)
        yield lib, ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _ops_ns(binding):
    """binding: "ops" (default) = torch.ops.holdem.* (csrc/torch_ops.cpp), "ctypes" = the plain C-ABI binding."""
    if binding == "ops":
        return _ops.load()
    if binding != "ctypes":
        raise ValueError("binding must be 'ops' or 'ctypes'")
    return None


# ---- device entry points ---------------------------------------------------------------------------------------------
def evaluate(cards: torch.Tensor, n_cards: int, out: torch.Tensor = None, strict: bool = False,
             binding: str = "ops") -> torch.Tensor:
    """Treys rank (1..7462, lower is better) of each row's first n_cards (5/6/7) cards.  cards: uint8 [N,8].
    Rows with a card > 51 give 0.  strict=True also gives 0 for rows with duplicate cards (the large-batch fast path
    does not look for duplicates, exactly like Treys)."""
    _require_cuda(cards, "cards", torch.uint8, 8)
    if n_cards not in (5, 6, 7):
        raise KeyError(n_cards)  # treys.Evaluator.hand_size_map
    ops = _ops_ns(binding)
    if ops is not None and out is None:
        return ops.eval(cards, n_cards, strict)
    if out is None:
        out = torch.empty(cards.shape[0], dtype=torch.int16, device=cards.device)
    with _enter(cards) as (lib, st):
        fn = lib.holdem_eval_batch_strict if strict else lib.holdem_eval_batch
        _lib.check(fn(cards.data_ptr(), n_cards, cards.shape[0], out.data_ptr(), st))
    return out


def category(cards: torch.Tensor, n_cards: int, binding: str = "ops") -> torch.Tensor:
    """Literal CalculateHandRankValue category 0..8 of each row's first n_cards (2..9) cards, duplicates allowed."""
    _require_cuda(cards, "cards", torch.uint8)
    if cards.dim() != 2 or cards.shape[1] < n_cards:
        raise ValueError("cards must be [N, >= n_cards]")
    ops = _ops_ns(binding)
    if ops is not None:
        return ops.category(cards, n_cards)
    out = torch.empty(cards.shape[0], dtype=torch.uint8, device=cards.device)
    with _enter(cards) as (lib, st):
        _lib.check(lib.holdem_category_batch(cards.data_ptr(), n_cards, cards.shape[1], cards.shape[0],
                                             out.data_ptr(), st))
    return out


def hand_strength_counts(sit: torch.Tensor, mode="treys", variant: int = None, binding: str = "ops") -> torch.Tensor:
    """(ahead, tied, behind) per situation: int32 [N,3].  sit: uint8 [N,8] rows h0 h1 b0..b4(0xFF) pad.
    variant (A/B measurements): 0 = production kernel, 1 = one evaluation per opponent holding."""
    _require_cuda(sit, "sit", torch.uint8, 8)
    ops = _ops_ns(binding)
    if ops is not None and variant is None:
        return ops.hs(sit, mode_id(mode))
    out = torch.empty((sit.shape[0], HS_ROW), dtype=torch.int32, device=sit.device)
    with _enter(sit) as (lib, st):
        if variant is None:
            _lib.check(lib.holdem_hs_batch(sit.data_ptr(), sit.shape[0], mode_id(mode), out.data_ptr(), st))
        else:
            _lib.check(lib.holdem_hs_batch_variant(sit.data_ptr(), sit.shape[0], mode_id(mode), int(variant),
                                                   out.data_ptr(), st))
    return out


def hand_potential_counts(sit: torch.Tensor, mode="treys", out: torch.Tensor = None,
                          direct: bool = False, flop_variant: int = None, turn_variant: int = None,
                          variants=None, ref_variant: int = None, binding: str = "ops") -> torch.Tensor:
    """HS + HP counts per situation: int32 [N,16] = ahead tied behind | HP[3][3] | HPTotal[3] | n_board.
    `out` may be a [N,16] int32 view into a larger buffer (e.g. this rank's slice of an all-gather buffer).
    direct=True forces the plain-enumeration kernel (the independent second implementation).  flop_variant (treys mode,
    flop rows only; 4 = rank-multiset kernel, 5 / 6 / 7 = that kernel without its closed-form add parts / in input order / with the
    step-by-step undo half, 3 / 2 = the 4-set kernels) runs one named implementation of the flop kernel
    for A/B measurements; rows of other streets are then left as they are in `out`.  turn_variant (4 = rank-multiset
    kernel, 3 = 3-set walk) does the same for turn rows.  variants=(flop, turn, river) runs a whole mixed-street treys batch
    through named kernels (holdem_hp_batch_variant).  ref_variant (mode='reference': 5 = rank-multiset kernel, 4 = the
    earlier 4-set kernel) does the same for the reference's literal mode."""
    _require_cuda(sit, "sit", torch.uint8, 8)
    plain = not direct and flop_variant is None and turn_variant is None and variants is None and ref_variant is None
    ops = _ops_ns(binding)
    if ops is not None and plain and out is None:
        return ops.hp(sit, mode_id(mode))
    if out is None:
        out = torch.empty((sit.shape[0], HP_ROW), dtype=torch.int32, device=sit.device)
    else:
        _require_cuda(out, "out", torch.int32, HP_ROW)
        if out.shape[0] != sit.shape[0] or out.device != sit.device:
            raise ValueError("out must be [N,16] on the same device as sit")
    if ops is not None and plain:
        ops.hp_out(sit, mode_id(mode), out)
        return out
    with _enter(sit) as (lib, st):
        if ref_variant is not None:
            if mode_id(mode) != MODE_REFERENCE or direct:
                raise ValueError("ref_variant applies to mode='reference' without direct=True")
            _lib.check(lib.holdem_hp_ref_variant(sit.data_ptr(), sit.shape[0], int(ref_variant), out.data_ptr(), st))
        elif variants is not None:
            if mode_id(mode) != MODE_TREYS or direct or flop_variant is not None or turn_variant is not None:
                raise ValueError("variants applies to mode='treys' alone")
            f, t, r = (int(v) for v in variants)
            _lib.check(lib.holdem_hp_batch_variant(sit.data_ptr(), sit.shape[0], f, t, r, out.data_ptr(), st))
        elif turn_variant is not None:
            if mode_id(mode) != MODE_TREYS or direct or flop_variant is not None:
                raise ValueError("turn_variant applies to mode='treys' alone")
            _lib.check(lib.holdem_hp_turn_variant(sit.data_ptr(), sit.shape[0], int(turn_variant), out.data_ptr(), st))
        elif flop_variant is not None:
            if mode_id(mode) != MODE_TREYS or direct:
                raise ValueError("flop_variant applies to mode='treys' without direct=True")
            _lib.check(lib.holdem_hp_flop_variant(sit.data_ptr(), sit.shape[0], int(flop_variant), out.data_ptr(), st))
        else:
            fn = lib.holdem_hp_batch_direct if direct else lib.holdem_hp_batch
            _lib.check(fn(sit.data_ptr(), sit.shape[0], mode_id(mode), out.data_ptr(), st))
    return out


def table_hand_potential_counts(hole: torch.Tensor, board: torch.Tensor, n_board: int, mode="treys",
                                binding: str = "ops") -> torch.Tensor:
    """Whole tables (GameBoard.py:12, :204-229): hole uint8 [H,S,2], board uint8 [H,5] -> int32 [H,S,16], the HS+HP rows
    of every seat on the first n_board (3/4/5) board cards."""
    _require_cuda(hole, "hole", torch.uint8)
    _require_cuda(board, "board", torch.uint8, 5)
    if hole.dim() != 3 or hole.shape[2] != 2 or board.shape[0] != hole.shape[0] or board.device != hole.device:
        raise ValueError("expected hole [H,S,2] and board [H,5] on one device")
    ops = _ops_ns(binding)
    if ops is not None:
        return ops.table_hp(hole, board, int(n_board), mode_id(mode))
    out = torch.empty((hole.shape[0], hole.shape[1], HP_ROW), dtype=torch.int32, device=hole.device)
    with _enter(hole) as (lib, st):
        _lib.check(lib.holdem_table_hp_batch(hole.data_ptr(), board.data_ptr(), hole.shape[0], hole.shape[1], int(n_board),
                                             mode_id(mode), out.data_ptr(), st))
    return out


def table_showdown(hole: torch.Tensor, board: torch.Tensor, active: torch.Tensor = None, binding: str = "ops"):
    """Showdown of whole tables in ONE kernel (GameBoard.determine_winner, GameBoard.py:204-229): hole uint8 [H,S,2], board
    uint8 [H,5], active optional [H,S] -> (ranks int16 [H,S], winners int32 [H,2]).  winners[:,0] is the seat bit mask of the
    best hand(s) among the active seats (lowest Treys rank), winners[:,1] of the seats the reference picks (it takes max() of
    the Treys values, GameBoard.py:210)."""
    _require_cuda(hole, "hole", torch.uint8)
    _require_cuda(board, "board", torch.uint8, 5)
    if hole.dim() != 3 or hole.shape[2] != 2 or board.shape[0] != hole.shape[0] or board.device != hole.device:
        raise ValueError("expected hole [H,S,2] and board [H,5] on one device")
    if active is not None and (not active.is_cuda or active.device != hole.device or tuple(active.shape) != tuple(hole.shape[:2])):
        raise ValueError("active must be [H,S] on the device of hole")
    ops = _ops_ns(binding)
    if ops is not None:
        return ops.table_showdown(hole, board, active)
    act = None if active is None else active.to(torch.uint8).contiguous()
    ranks = torch.empty(tuple(hole.shape[:2]), dtype=torch.int16, device=hole.device)
    win = torch.empty((hole.shape[0], 2), dtype=torch.int32, device=hole.device)
    with _enter(hole) as (lib, st):
        _lib.check(lib.holdem_table_showdown(hole.data_ptr(), board.data_ptr(), 0 if act is None else act.data_ptr(),
                                             hole.shape[0], hole.shape[1], ranks.data_ptr(), win.data_ptr(), st))
    return ranks, win


def flop_phase_cycles(sit: torch.Tensor):
    """Production flop kernel with phase clocks on -> (rows int32 [N,16], cycles int64 [8]; SM cycles of one thread per situation
    group summed over all situations: [0] card lists + table rows, [1] hand strength + rank-pair sets, [4] our-flush entries,
    [5] step descriptors, [2] next claim + wait for the T row, [6] generic tasks up to the group barrier, [3] flush tasks + reduce,
    [7] unused)."""
    _require_cuda(sit, "sit", torch.uint8, 8)
    out = torch.empty((sit.shape[0], HP_ROW), dtype=torch.int32, device=sit.device)
    cyc = torch.zeros(8, dtype=torch.int64, device=sit.device)
    with _enter(sit) as (lib, st):
        _lib.check(lib.holdem_hp_flop_phase_cycles(sit.data_ptr(), sit.shape[0], out.data_ptr(), cyc.data_ptr(), st))
    return out, cyc


def exhaustive7(device=None, variant: int = None):
    """All C(52,7) hands evaluated in-kernel -> (hist int64 [7463] indexed by rank, sums int64 [2] = sum r, sum r^2).
    variant: None / 0 = tuple kernel (production), 1 = chunk-per-thread kernel."""
    dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
    hist = torch.empty(7463, dtype=torch.int64, device=dev)
    sums = torch.empty(2, dtype=torch.int64, device=dev)
    with _enter(hist) as (lib, st):
        if variant is None:
            _lib.check(lib.holdem_exhaustive7(hist.data_ptr(), sums.data_ptr(), st))
        else:
            _lib.check(lib.holdem_exhaustive7_variant(int(variant), hist.data_ptr(), sums.data_ptr(), st))
    return hist, sums


def rows_hash(rows: torch.Tensor, first_row: int = 0) -> int:
    """64-bit content hash of a contiguous int32 [n, words] CUDA buffer (holdem_rows_hash): position-dependent through
    first_row + row index, additive over shards (mod 2^64).  Synchronises (returns a Python int)."""
    _require_cuda(rows, "rows", torch.int32)
    if rows.dim() != 2:
        raise ValueError("rows must be [n, words]")
    out = torch.empty(1, dtype=torch.int64, device=rows.device)
    with _enter(rows) as (lib, st):
        _lib.check(lib.holdem_rows_hash(rows.data_ptr(), rows.shape[0], rows.shape[1], int(first_row), out.data_ptr(), st))
    return int(out.item()) & 0xFFFFFFFFFFFFFFFF


def rows_hash_host(rows, first_row: int = 0) -> int:
    """numpy twin of rows_hash (same value) for host-side checks of buffers that never were on a device."""
    r = np.ascontiguousarray(rows).view(np.uint32).astype(np.uint64)
    n = r.shape[0]
    M = np.uint64

    def mix(x):
        x = x ^ (x >> M(30)); x = x * M(0xBF58476D1CE4E5B9)
        x = x ^ (x >> M(27)); x = x * M(0x94D049BB133111EB)
        return x ^ (x >> M(31))

    with np.errstate(over="ignore"):
        h = M(0xCBF29CE484222325) ^ mix(np.arange(first_row, first_row + n, dtype=np.uint64) + M(0x9E3779B97F4A7C15))
        for w in range(r.shape[1]):
            h = (h ^ r[:, w]) * M(0x100000001B3)
        return int(mix(h).sum(dtype=np.uint64))


def pack_rank_major(hole: torch.Tensor, board: torch.Tensor) -> torch.Tensor:
    """utils.card_to_int wire ints (utils.py:82-85: rank * 4 + suit; int64 [N,2] and [N,3..5], CUDA) -> uint8 [N,8]
    situation rows in the path encoding.  Values outside [0,52) become 0xFE, which the kernels report as malformed rows."""
    return _ops.load().pack_rank_major(hole.contiguous(), board.contiguous())


def kernel_launches() -> int:
    """Kernels launched by the library in this process so far (every launch site counts itself)."""
    return int(_lib.load().holdem_kernel_launches())


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
@contextlib.contextmanager
def _host_enter(device):
    dev = torch.cuda.current_device() if device is None else int(device)
    with torch.cuda.device(dev):
        yield _lib.load() if dev in _lib._inited else _lib.init(dev)


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
    with _host_enter(device) as lib:
        _lib.check(lib.holdem_eval_batch_host(cards.ctypes.data, n_cards, cards.shape[0], out.ctypes.data))
    return out


def category_host(cards, n_cards: int, device=None) -> np.ndarray:
    cards = _np_u8(cards)
    out = np.empty(cards.shape[0], dtype=np.uint8)
    with _host_enter(device) as lib:
        _lib.check(lib.holdem_category_batch_host(cards.ctypes.data, n_cards, cards.shape[1], cards.shape[0],
                                                  out.ctypes.data))
    return out


def hand_strength_counts_host(sit, mode="treys", device=None) -> np.ndarray:
    sit = _np_u8(sit, 8)
    out = np.empty((sit.shape[0], HS_ROW), dtype=np.uint32)
    with _host_enter(device) as lib:
        _lib.check(lib.holdem_hs_batch_host(sit.ctypes.data, sit.shape[0], mode_id(mode), out.ctypes.data))
    return out


def hand_potential_counts_host(sit, mode="treys", device=None, out=None, staged: bool = False) -> np.ndarray:
    """staged=True: the chunked copy pipeline even for page-locked buffers (holdem_hp_batch_host_staged)."""
    sit = _np_u8(sit, 8)
    if out is None:
        out = np.empty((sit.shape[0], HP_ROW), dtype=np.uint32)
    with _host_enter(device) as lib:
        fn = lib.holdem_hp_batch_host_staged if staged else lib.holdem_hp_batch_host
        _lib.check(fn(sit.ctypes.data, sit.shape[0], mode_id(mode), out.ctypes.data))
    return out


def smem_gather_microbench(table_bytes=131072, iters=2048, pattern=0, blocks_per_sm=1, threads=1024, device=None):
    """Measured shared-memory uint16 gather rate (lookups/s) -- the lookup roofline denominator."""
    ms = ctypes.c_float(0)
    lookups = ctypes.c_double(0)
    with _host_enter(device) as lib:
        _lib.check(lib.holdem_microbench_smem_gather(table_bytes, iters, pattern, blocks_per_sm, threads,
                                                     ctypes.byref(ms), ctypes.byref(lookups), None))
    return lookups.value / (ms.value * 1e-3), ms.value


def int_issue_microbench(mix=0, iters=4096, blocks_per_sm=1, threads=1024, device=None):
    """Measured integer instruction issue rate (warp instructions/s) -- the issue roofline denominator.
    mix 0: compare + predicated multiply-add pairs (ALU + FMA pipes), 1: FMA pipe only, 2: ALU pipe only."""
    ms = ctypes.c_float(0)
    n = ctypes.c_double(0)
    with _host_enter(device) as lib:
        _lib.check(lib.holdem_microbench_int_issue(iters, mix, blocks_per_sm, threads, ctypes.byref(ms), ctypes.byref(n), None))
    return n.value / (ms.value * 1e-3), ms.value
