"""Batched mirror of the reference's Decision.decision_based_on_blind_position (Decision.py:4-48), SURVEY.md 8f2.

The reference evaluates one (hand_strength, ppot, npot, position) tuple at a time in Python; this is the same threshold
tree applied to whole batches.  CUDA tensors go through ONE fused elementwise kernel (csrc/kernels_aux.cu:
holdem_decide_batch, and holdem_decide_from_counts which derives HS and the normalised potentials from the count rows in
the same pass); CPU tensors (the scalar drop-in form and the CPU tests) use the equivalent torch expressions below.  The
strings and thresholds are the reference's.  Note the reference returns None (no branch taken) when hand_strength > 0.8
and npot >= 0.3; that case is code -1 / None here too.

The thresholds only make sense for potentials that are probabilities: feed `normalised_potentials()`, not the raw
PPot/NPot of the reference formula.
"""
import torch

ACTIONS = ("fold", "check", "call", "raise")
FOLD, CHECK, CALL, RAISE, NONE_TAKEN = 0, 1, 2, 3, -1
POSITIONS = {"SB": 0, "BB": 1, "NONE": 2}


def position_code(position) -> int:
    if position not in POSITIONS:
        raise ValueError("Invalid position. Expected 'SB', 'BB', or 'NONE'.")
    return POSITIONS[position]


def position_from_player_position(player_position) -> str:
    """utils.PlayerPosition (utils.py:45-50: NONE, SMALL_BLIND, BIG_BLIND, DEALER, DEALER_SMALL_BLIND) -> the 'SB' / 'BB' /
    'NONE' strings Decision.py expects.  Accepts the Enum member, its name or its auto() value (1..5).  The heads-up
    DEALER_SMALL_BLIND posts the small blind; DEALER posts nothing."""
    name = getattr(player_position, "name", player_position)
    if isinstance(name, int):
        name = {1: "NONE", 2: "SMALL_BLIND", 3: "BIG_BLIND", 4: "DEALER", 5: "DEALER_SMALL_BLIND"}.get(name)
    table = {"NONE": "NONE", "DEALER": "NONE", "SMALL_BLIND": "SB", "DEALER_SMALL_BLIND": "SB", "BIG_BLIND": "BB"}
    if name not in table:
        raise ValueError(f"not a PlayerPosition: {player_position!r}")
    return table[name]


def position_codes(positions, device=None) -> torch.Tensor:
    """Iterable of PlayerPosition members / names / 'SB' 'BB' 'NONE' strings -> uint8 tensor of position codes."""
    codes = [POSITIONS[p] if p in POSITIONS else POSITIONS[position_from_player_position(p)] for p in positions]
    return torch.tensor(codes, dtype=torch.uint8, device=device)


def _pos_args(position, shape, device):
    """-> (per-row uint8 tensor or None, code for the whole batch)."""
    if isinstance(position, str):
        return None, position_code(position)
    pos = position.to(device=device)
    if bool(((pos < 0) | (pos > 2)).any()):
        raise ValueError("Invalid position. Expected 'SB', 'BB', or 'NONE'.")
    return pos.to(torch.uint8).reshape(-1).contiguous(), 2


def decide_batch(hand_strength: torch.Tensor, ppot: torch.Tensor, npot: torch.Tensor, position) -> torch.Tensor:
    """-> int8 tensor of action codes (FOLD/CHECK/CALL/RAISE, -1 where the reference falls through and returns None).
    `position` is 'SB' / 'BB' / 'NONE' for the whole batch or an integer tensor of position codes (0/1/2)."""
    hs, pp, np_ = hand_strength, ppot, npot
    if hs.is_cuda:
        from . import _ops
        pos, pos_all = _pos_args(position, hs.shape, hs.device)
        f64 = lambda t: t.to(torch.float64).contiguous()
        return _ops.load().decide(f64(hs), f64(pp), f64(np_), pos, pos_all)
    if isinstance(position, str):
        pos = torch.full(hs.shape, position_code(position), dtype=torch.int8, device=hs.device)
    else:
        pos = position.to(device=hs.device, dtype=torch.int8)
        if bool(((pos < 0) | (pos > 2)).any()):
            raise ValueError("Invalid position. Expected 'SB', 'BB', or 'NONE'.")
    sb, bb = pos == 0, pos == 1

    def pick(a_sb, a_bb, a_none):
        out = torch.full(hs.shape, a_none, dtype=torch.int8, device=hs.device)
        out = torch.where(bb, torch.full_like(out, a_bb), out)
        return torch.where(sb, torch.full_like(out, a_sb), out)

    high = hs > 0.8
    mid = (~high) & (hs > 0.5)
    res_high = torch.where(np_ < 0.3, pick(CHECK, CALL, RAISE), torch.full(hs.shape, NONE_TAKEN, dtype=torch.int8,
                                                                           device=hs.device))
    res_mid = torch.where(pp > 0.6, pick(CHECK, CALL, CALL), pick(FOLD, CHECK, FOLD))
    low_bb = torch.where(pp > 0.8, torch.full(hs.shape, CALL, dtype=torch.int8, device=hs.device),
                         torch.full(hs.shape, CHECK, dtype=torch.int8, device=hs.device))
    res_low = torch.where(bb, low_bb, torch.full(hs.shape, FOLD, dtype=torch.int8, device=hs.device))
    return torch.where(high, res_high, torch.where(mid, res_mid, res_low))


def decision_based_on_blind_position(hand_strength, ppot, npot, position):
    """Scalar form with the reference's signature and return values ('check' / 'call' / 'raise' / 'fold' / None)."""
    code = int(decide_batch(torch.tensor([float(hand_strength)], dtype=torch.float64),
                            torch.tensor([float(ppot)], dtype=torch.float64),
                            torch.tensor([float(npot)], dtype=torch.float64), position)[0])
    return None if code == NONE_TAKEN else ACTIONS[code]


def decide_from_counts(rows: torch.Tensor, position) -> torch.Tensor:
    """HS/HP count rows (int32 [N,16] from hand_potential_counts) -> action codes, using HS and the NORMALISED
    potentials."""
    if rows.is_cuda:
        from . import _ops
        pos, pos_all = _pos_args(position, rows.shape[:-1], rows.device)
        return _ops.load().decide_from_counts(rows.reshape(-1, 16).contiguous(), pos, pos_all).reshape(rows.shape[:-1])
    from .api import hs_from_counts, normalised_potentials
    hs = hs_from_counts(rows)
    pp, np_ = normalised_potentials(rows)
    return decide_batch(hs, torch.nan_to_num(pp, nan=0.0), torch.nan_to_num(np_, nan=0.0), position)
