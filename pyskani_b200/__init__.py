"""pyskani_b200 — pyskani's API (`Database`, `Sketch`, `Hit`) running skani's sketch -> screen -> chain -> ANI path
on NVIDIA B200 GPUs.

Drop-in for the reference package's public surface (src/pyskani/__init__.py:1-31):

    >>> import pyskani_b200 as pyskani
    >>> db = pyskani.Database()
    >>> db.sketch("ref", ref_sequence)
    >>> db.query("query", query_sequence)        # -> list of Hit

The compiled extension `_skani` calls libskb.so (hand-written sm_100a CUDA kernels) through its C ABI
(include/skb.h).  There is no CPU fallback: using a Database without a CUDA device raises RuntimeError.
"""
from . import _skani
from ._skani import Sketch, Database, Hit

__version__ = _skani.__version__
__author__ = _skani.__author__
__doc__ = __doc__
__build__ = _skani.__build__
__all__ = [
    "Sketch",
    "Database",
    "Hit",
    "SKANI_VERSION",
]

# Version of the skani algorithm the kernels restate (the reference exposes the embedded crate's version)
SKANI_VERSION = _skani.__build__["dependencies"]["skani"]
