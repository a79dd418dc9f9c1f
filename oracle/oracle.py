"""ctypes front-end of the CPU oracle (oracle/treys_oracle.c, oracle/ehs_oracle.c).

TEST INFRASTRUCTURE ONLY: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.  The product
(dask-holdem-simulator_b200/) never imports anything from oracle/.

Float formulas restate HandCalculations.py:175 (Python float / double) and
HandCalculations.py:221-224 (float32 tensor arithmetic, this exact operation order).
"""
import ctypes
import os
import subprocess
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None
_lock = threading.Lock()

MODE_TREYS = 0
MODE_REFERENCE = 1


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("treys_oracle.c", "ehs_oracle.c")]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _LIB_PATH


def lib():
    global _lib
    with _lock:
        if _lib is None:
            build()
            L = ctypes.CDLL(_LIB_PATH)
            i32p = ctypes.POINTER(ctypes.c_int)
            u32p = ctypes.POINTER(ctypes.c_uint32)
            u8p = ctypes.POINTER(ctypes.c_uint8)
            u16p = ctypes.POINTER(ctypes.c_uint16)
            u64p = ctypes.POINTER(ctypes.c_uint64)
            L.treys_init.restype = None
            L.treys_flush_keys.restype = ctypes.c_int
            L.treys_unsuited_keys.restype = ctypes.c_int
            L.treys_card_new.argtypes = [ctypes.c_int, ctypes.c_char]
            L.treys_card_new.restype = ctypes.c_uint32
            L.treys_card_from_path.argtypes = [ctypes.c_int]
            L.treys_card_from_path.restype = ctypes.c_uint32
            L.treys_evaluate.argtypes = [i32p, ctypes.c_int]
            L.treys_evaluate.restype = ctypes.c_int
            L.treys_evaluate_batch.argtypes = [u8p, ctypes.c_int, ctypes.c_int64, u16p]
            L.treys_evaluate_batch.restype = None
            L.treys_exhaustive7.argtypes = [ctypes.c_int, ctypes.c_int, u64p, u64p, u64p]
            L.treys_exhaustive7.restype = None
            L.lit_is_straight.argtypes = [i32p, ctypes.c_int]
            L.lit_is_straight.restype = ctypes.c_int
            L.lit_category.argtypes = [i32p, ctypes.c_int]
            L.lit_category.restype = ctypes.c_int
            L.ehs_hs_counts.argtypes = [ctypes.c_int, i32p, i32p, ctypes.c_int, u32p]
            L.ehs_hs_counts.restype = None
            L.ehs_hp_counts.argtypes = [ctypes.c_int, i32p, i32p, ctypes.c_int, u32p, u32p]
            L.ehs_hp_counts.restype = None
            L.ehs_hp_counts_treys_cached.argtypes = [i32p, i32p, ctypes.c_int, u32p, u32p]
            L.ehs_hp_counts_treys_cached.restype = None
            L.ehs_hp_counts_treys_4set.argtypes = [i32p, i32p, u32p, u32p]
            L.ehs_hp_counts_treys_4set.restype = None
            L.ehs_hp_batch_ex.argtypes = [ctypes.c_int, ctypes.c_int, u8p, ctypes.c_int64, u32p]
            L.ehs_hp_batch_ex.restype = None
            L.ehs_hs_batch.argtypes = [ctypes.c_int, u8p, ctypes.c_int64, u32p]
            L.ehs_hs_batch.restype = None
            L.ehs_hp_batch.argtypes = [ctypes.c_int, u8p, ctypes.c_int64, u32p]
            L.ehs_hp_batch.restype = None
            L.lit_category_batch.argtypes = [u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int64, u8p]
            L.lit_category_batch.restype = None
            L.treys_init()
            _lib = L
    return _lib


def _ints(seq):
    seq = [int(x) for x in seq]
    return (ctypes.c_int * len(seq))(*seq), len(seq)


def _mode(mode) -> int:
    if mode in (MODE_TREYS, "treys"):
        return MODE_TREYS
    if mode in (MODE_REFERENCE, "reference"):
        return MODE_REFERENCE
    raise ValueError(f"unknown mode {mode!r}")


# ---- Treys ---------------------------------------------------------------------------------
def treys_evaluate(cards) -> int:
    """Treys rank 1..7462 of 5/6/7 path-encoded cards (rank = c % 13, suit = c // 13)."""
    arr, n = _ints(cards)
    r = lib().treys_evaluate(arr, n)
    if r < 0:
        raise KeyError(n)  # Evaluator.hand_size_map[len(all_cards)]
    return r


def treys_evaluate_batch(rows: np.ndarray, n_cards: int) -> np.ndarray:
    rows = np.ascontiguousarray(rows, dtype=np.uint8)
    assert rows.ndim == 2 and rows.shape[1] == 8
    out = np.empty(rows.shape[0], dtype=np.uint16)
    lib().treys_evaluate_batch(rows.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), n_cards,
                               rows.shape[0], out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint16)))
    return out


def treys_card_from_path(card: int) -> int:
    return int(lib().treys_card_from_path(int(card)))


def treys_exhaustive7(c0_lo=0, c0_hi=46):
    hist = np.zeros(7463, dtype=np.uint64)
    s = ctypes.c_uint64(0)
    s2 = ctypes.c_uint64(0)
    lib().treys_exhaustive7(c0_lo, c0_hi, hist.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)),
                            ctypes.byref(s), ctypes.byref(s2))
    return hist, int(s.value), int(s2.value)


# ---- literal ranker ----------------------------------------------------------------------------
def lit_is_straight(values) -> bool:
    arr, n = _ints(values)
    return bool(lib().lit_is_straight(arr, n))


def lit_category(cards) -> int:
    arr, n = _ints(cards)
    return int(lib().lit_category(arr, n))


def lit_category_batch(rows: np.ndarray, n_cards: int) -> np.ndarray:
    rows = np.ascontiguousarray(rows, dtype=np.uint8)
    out = np.empty(rows.shape[0], dtype=np.uint8)
    lib().lit_category_batch(rows.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), n_cards,
                             rows.shape[1], rows.shape[0], out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)))
    return out


# ---- HS / HP counts --------------------------------------------------------------------------------
def hs_counts(our, board, mode):
    o, _ = _ints(our)
    b, nb = _ints(board)
    out = (ctypes.c_uint32 * 3)()
    lib().ehs_hs_counts(_mode(mode), o, b, nb, out)
    return tuple(int(v) for v in out)


def hp_counts(our, board, mode, cached=True):
    """-> (HP 3x3 list of ints, HPTotal list of 3 ints).  cached: False = the plain loop, True = our rank cached per
    runout, "4set" = also the opponent's rank cached per 4-set of unseen cards (treys mode, flop only)."""
    o, _ = _ints(our)
    b, nb = _ints(board)
    hp = (ctypes.c_uint32 * 9)()
    tot = (ctypes.c_uint32 * 3)()
    m = _mode(mode)
    if m == MODE_TREYS and cached == "4set":
        assert nb == 3
        lib().ehs_hp_counts_treys_4set(o, b, hp, tot)
    elif m == MODE_TREYS and cached:
        lib().ehs_hp_counts_treys_cached(o, b, nb, hp, tot)
    else:
        lib().ehs_hp_counts(m, o, b, nb, hp, tot)
    v = [int(x) for x in hp]
    return [v[0:3], v[3:6], v[6:9]], [int(x) for x in tot]


def pack_situations(situations) -> np.ndarray:
    """[(our[2], board[3..5]), ...] -> uint8 [N,8]: h0 h1 b0..b4 (0xFF absent) pad(0xFF)."""
    rows = np.full((len(situations), 8), 0xFF, dtype=np.uint8)
    for i, (our, board) in enumerate(situations):
        rows[i, 0:2] = our
        rows[i, 2:2 + len(board)] = board
    return rows


def hs_batch(sit: np.ndarray, mode) -> np.ndarray:
    sit = np.ascontiguousarray(sit, dtype=np.uint8)
    out = np.zeros((sit.shape[0], 3), dtype=np.uint32)
    lib().ehs_hs_batch(_mode(mode), sit.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), sit.shape[0],
                       out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)))
    return out


def hp_batch(sit: np.ndarray, mode, cache4: bool = False) -> np.ndarray:
    """-> uint32 [N,16]: ahead tied behind | HP[9] | HPTotal[3] | n_board.
    cache4=True: treys-mode flop rows use the 4-set cache (~6x fewer evaluations; same counts)."""
    sit = np.ascontiguousarray(sit, dtype=np.uint8)
    out = np.zeros((sit.shape[0], 16), dtype=np.uint32)
    lib().ehs_hp_batch_ex(_mode(mode), int(bool(cache4)), sit.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)),
                          sit.shape[0], out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)))
    return out


def _hp_chunk(args):
    sit_bytes, n, mode, cache4 = args
    sit = np.frombuffer(sit_bytes, dtype=np.uint8).reshape(n, 8)
    return hp_batch(sit, mode, cache4).tobytes()


def hp_batch_parallel(sit: np.ndarray, mode, cache4: bool = True, procs: int = None) -> np.ndarray:
    """hp_batch over all host cores (one forked process per core, rows dealt round-robin)."""
    import multiprocessing as mp
    sit = np.ascontiguousarray(sit, dtype=np.uint8)
    n = sit.shape[0]
    if procs is None:
        try:
            procs = len(os.sched_getaffinity(0))
        except Exception:
            procs = os.cpu_count() or 1
    procs = max(1, min(procs, n))
    lib()
    out = np.zeros((n, 16), dtype=np.uint32)
    if procs == 1:
        return hp_batch(sit, mode, cache4)
    parts_in = [np.ascontiguousarray(sit[i::procs]) for i in range(procs)]
    with mp.get_context("fork").Pool(procs) as pool:
        parts = pool.map(_hp_chunk, [(p.tobytes(), len(p), mode, cache4) for p in parts_in])
    for i, p in enumerate(parts):
        out[i::procs] = np.frombuffer(p, dtype=np.uint32).reshape(-1, 16)
    return out


# ---- float formulas -----------------------------------------------------------------------------------
def hs_float(ahead: int, tied: int, behind: int) -> float:
    """HandCalculations.py:175 -- Python float arithmetic (double)."""
    return (ahead + tied / 2) / (ahead + tied + behind)


def hp_floats(HP, HPTotal):
    """HandCalculations.py:221-224 -- float32, the reference's operation order; x/0 -> inf, 0/0 -> nan."""
    f = np.float32
    hp = [[f(v) for v in row] for row in HP]
    t = [f(v) for v in HPTotal]
    ahead, tied, behind = 0, 1, 2
    with np.errstate(divide="ignore", invalid="ignore"):
        ppot = (hp[behind][ahead] + hp[behind][tied] / f(2) + hp[tied][ahead] / f(2)) / (t[behind] + t[tied])
        npot = (hp[ahead][behind] + hp[tied][behind] / f(2) + hp[ahead][tied] / f(2)) / (t[ahead] + t[tied])
    return f(ppot), f(npot)
