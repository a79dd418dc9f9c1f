"""Import the UNMODIFIED reference `HandCalculations.py` behind stub modules.

TEST INFRASTRUCTURE ONLY (container-only: /root/reference does not exist on the
GPU box).  Used by tests/golden/make_golden.py to generate the committed golden
vectors and by the CPU tests to re-validate the C restatement when the
reference tree is present.

The reference (HandCalculations.py:1-9) imports `utils` (needs poker, colorama),
`treys` and `pokerlib_to_treys` (needs poker, treys).  None of those is touched
by the hot path (HandCalculations.py:24-82, 131-226), so minimal fakes in
sys.modules are enough to import the file exactly as shipped.
"""
import enum
import os
import sys
import types

REFERENCE_DIR = os.environ.get("HOLDEM_REFERENCE_DIR", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "HandCalculations.py"))


def _install_stubs() -> None:
    if "poker" not in sys.modules:
        poker = types.ModuleType("poker")
        # member names needed as dict keys at utils.py:15-36
        poker.Rank = enum.Enum(
            "Rank", "DEUCE THREE FOUR FIVE SIX SEVEN EIGHT NINE TEN JACK QUEEN KING ACE")
        poker.Suit = enum.Enum("Suit", "CLUBS DIAMONDS HEARTS SPADES")

        class Card:  # never constructed on the hot path
            def __init__(self, *a, **k):
                self.args = a

        poker.Card = Card
        sys.modules["poker"] = poker
    if "colorama" not in sys.modules:
        colorama = types.ModuleType("colorama")

        class _Blank:
            def __getattr__(self, name):
                return ""

        colorama.Fore = _Blank()
        colorama.Style = _Blank()
        colorama.init = lambda **kw: None
        sys.modules["colorama"] = colorama
    if "treys" not in sys.modules:
        treys = types.ModuleType("treys")

        class Evaluator:  # constructed at import (HandCalculations.py:268)
            def evaluate(self, *a, **k):
                raise RuntimeError("treys is not installed; stub Evaluator")

        class TCard:
            @staticmethod
            def new(s):
                raise RuntimeError("treys is not installed; stub Card")

        treys.Evaluator = Evaluator
        treys.Card = TCard
        sys.modules["treys"] = treys
    if "timeout_decorator" not in sys.modules:
        td = types.ModuleType("timeout_decorator")
        td.timeout = lambda *a, **k: (lambda f: f)
        sys.modules["timeout_decorator"] = td


def load_reference():
    """Return the reference's HandCalculations module (CPU, unchanged)."""
    if not reference_available():
        raise FileNotFoundError(f"reference tree not found at {REFERENCE_DIR}")
    _install_stubs()
    # Make sure OUR drop-in of the same name is not picked up instead.
    for name in ("HandCalculations", "utils", "pokerlib_to_treys"):
        mod = sys.modules.get(name)
        if mod is not None and not str(getattr(mod, "__file__", "")).startswith(REFERENCE_DIR):
            del sys.modules[name]
    sys.path.insert(0, REFERENCE_DIR)
    try:
        import HandCalculations as ref  # noqa: the reference's file
    finally:
        sys.path.remove(REFERENCE_DIR)
    assert ref.__file__.startswith(REFERENCE_DIR), ref.__file__
    # The reference picks `device` at import (HandCalculations.py:22) and reads it at call time: pin it to the CPU here instead
    # of hiding the GPUs from the whole process through os.environ (a GPU test running later in the same session would lose them).
    ref.device = "cpu"
    # keep the reference's helper modules from shadowing ours afterwards
    for name in ("HandCalculations", "utils", "pokerlib_to_treys"):
        sys.modules.pop(name, None)
    return ref


class TorchZerosSpy:
    """Proxy for the `torch` name inside the reference module that records the
    tensors created by torch.zeros (HP and HPTotal, HandCalculations.py:187-188)
    so the integer count matrices can be frozen into the golden fixtures without
    editing the reference."""

    def __init__(self, torch_mod):
        self._torch = torch_mod
        self.captured = []

    def zeros(self, *a, **k):
        t = self._torch.zeros(*a, **k)
        self.captured.append(t)
        return t

    def __getattr__(self, name):
        return getattr(self._torch, name)
