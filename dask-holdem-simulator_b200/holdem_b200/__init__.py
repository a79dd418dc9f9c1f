"""holdem_b200 -- B200-native hand-strength / hand-potential path of dask-holdem-simulator.

Python mirror of the reference's interface for this path lives one directory up (HandCalculations.py,
pokerlib_to_treys.py); this package holds the C-ABI binding and the batched / sharded entry points.
"""
from . import _lib, _ops  # noqa: F401
from ._lib import HoldemError, LIB_PATH, HP_ROW, HS_ROW, MODE_TREYS, MODE_REFERENCE  # noqa: F401
from .api import *  # noqa: F401,F403
from .encoding import (pack_situations, random_situations, path_from_str, str_from_path,  # noqa: F401
                       treys_from_path, path_from_treys, treys_from_str, path_from_poker_card,
                       path_from_rank_major, rank_major_from_path, path_from_rank_major_t, rank_major_from_path_t,
                       treys_from_path_t, path_from_treys_t, situations_from_rank_major)
from .sharded import shard_bounds, hand_potential_counts_sharded  # noqa: F401
from .extensions import hand_strength_vs_n, preflop_hand_strength, monte_carlo_equity  # noqa: F401
from . import policy, table, extensions  # noqa: F401,E402

try:        # register torch.ops.holdem.* now when the extension is built; a missing build raises at first use instead
    _ops.load()
except RuntimeError:
    pass
