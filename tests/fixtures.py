"""Shared test helpers: golden E. coli pair and synthetic genome generators."""
import os

import numpy as np

_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ecoli_pair.npz")
_cache = {}


def unpack2(packed, n):
    codes = np.empty((len(packed), 4), np.uint8)
    for i in range(4):
        codes[:, i] = (packed >> (2 * i)) & 3
    return np.frombuffer(b"ACGT", np.uint8)[codes.reshape(-1)[:n]].tobytes()


def ecoli_pair():
    """(EC590 bytes, K12 bytes, goldens dict) — reference src/pyskani/tests/test_ani.py fixtures."""
    if "pair" not in _cache:
        z = np.load(_GOLDEN)
        ec = unpack2(z["EC590_packed"], int(z["EC590_len"]))
        k12 = unpack2(z["K12_packed"], int(z["K12_len"]))
        gold = {k[len("golden_"):]: float(z[k]) for k in z.files if k.startswith("golden_")}
        _cache["pair"] = (ec, k12, gold)
    return _cache["pair"]
