#!/usr/bin/env python
"""Golden vectors for the batched Decision mirror: runs the UNMODIFIED reference Decision.py (pure Python, no
third-party imports) over a grid of inputs.  Build container only."""
import itertools
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
import Decision  # noqa: E402  the reference's file

vals = [0.0, 0.25, 0.3, 0.3000001, 0.5, 0.5000001, 0.55, 0.6, 0.6000001, 0.7, 0.8, 0.8000001, 0.9, 1.0]
cases = []
for hs, pp, npot, pos in itertools.product(vals, vals, [0.0, 0.29, 0.3, 0.31, 1.0], ["SB", "BB", "NONE"]):
    cases.append([hs, pp, npot, pos, Decision.decision_based_on_blind_position(hs, pp, npot, pos)])
with open(os.path.join(HERE, "decision_golden.json"), "w") as f:
    json.dump({"source": "reference Decision.py:4-48 (unmodified)", "cases": cases}, f)
print(len(cases), {a: sum(1 for c in cases if c[4] == a) for a in ("check", "call", "raise", "fold", None)})
