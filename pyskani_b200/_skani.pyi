"""Type stubs of the C++ extension ``pyskani_b200._skani`` (source: ``pyskani_b200/ext/skani_module.cpp``).

The three classes carry the call signatures of pyskani's ``Database`` / ``Sketch`` / ``Hit`` (reference
``src/pyskani/_skani/lib.rs`` and ``hit.rs``) so that ``import pyskani_b200 as pyskani`` is a drop-in; everything
behind them runs on the GPU through ``libskb.so`` (``include/skb.h``).  Behavioural differences are listed in README.md.
"""
from __future__ import annotations

import os
from array import array
from pathlib import Path
from types import TracebackType
from typing import Literal, Sequence as _Seq, Tuple

StorageFormat = Literal["consolidated", "separated"]
PathLike = str | bytes | os.PathLike[str]
Contig = str | bytes | bytearray | memoryview | array
NamedGenome = Tuple[str, _Seq[Contig]]          # (name, contigs) as taken by sketch_many / query_many

__version__: str
__author__: str
__build__: dict[str, object]                     # {"cuda_arch": "sm_100a", "libskb": <skb_version()>, ...}


class Database:
    """An ordered set of sketches resident in HBM plus, optionally, a folder on disk (lib.rs:132-137).

    ``path=None`` keeps everything in memory; with a path, ``format`` selects one file per sketch (``"separated"``:
    ``markers.bin`` + ``<name>.sketch``) or two files (``"consolidated"``: ``sketches.db`` + ``index.db``).
    """

    def __init__(self, path: PathLike | None = None, *, compression: int = 125, marker_compression: int = 1000,
                 k: int = 15, format: StorageFormat | None = None) -> None: ...

    # -- construction from disk (lib.rs:423-470): both load every sketch into device memory; ``open`` keeps the folder
    #    attached so that later ``sketch()`` calls append to it
    @classmethod
    def open(cls, path: PathLike) -> Database: ...
    @classmethod
    def load(cls, path: PathLike) -> Database: ...

    # -- the hot path
    def sketch(self, name: str, *contigs: Contig, seed: bool = True) -> None:
        """FracMinHash-sketch one genome on the GPU and add it (lib.rs:472-512). Contigs shorter than 500 bp are skipped."""
    def query(self, name: str, *contigs: Contig, seed: bool = True, learned_ani: bool | None = None,
              median: bool = False, robust: bool = False, cutoff: float | None = None,
              faster_small: bool = False) -> list[Hit]:
        """Sketch the query, screen it against every reference, chain the survivors, return hits with ANI > 0.1
        (lib.rs:549-660).  ``learned_ani=True`` raises: the regression model is not available (README.md)."""
    def sketch_many(self, items: _Seq[NamedGenome], *, seed: bool = True) -> None:
        """``sketch`` for many genomes in one device batch (no counterpart in the reference)."""
    def query_many(self, items: _Seq[NamedGenome], *, seed: bool = True, learned_ani: bool | None = None,
                   median: bool = False, robust: bool = False, cutoff: float | None = None,
                   faster_small: bool = False) -> list[list[Hit]]:
        """``query`` for many genomes at once; one hit list per item, in order."""

    # -- persistence (lib.rs:662-744)
    def save(self, path: PathLike, overwrite: bool = False, format: StorageFormat | None = None) -> None: ...
    def flush(self) -> None: ...

    # -- introspection
    def __len__(self) -> int: ...
    @property
    def compression(self) -> int: ...
    @property
    def marker_compression(self) -> int: ...
    @property
    def path(self) -> Path | None: ...

    def __enter__(self) -> Database: ...
    def __exit__(self, exc_type: type[BaseException] | None, exc: BaseException | None,
                 traceback: TracebackType | None) -> bool | None: ...


class Hit:
    """One (query, reference) pair that survived screen, chaining and the ``ani > 0.1`` gate (hit.rs:19-122)."""

    def __init__(self, identity: float, query_name: str, query_fraction: float, reference_name: str,
                 reference_fraction: float) -> None: ...
    def __repr__(self) -> str: ...
    @property
    def reference_name(self) -> str: ...
    @property
    def query_name(self) -> str: ...
    @property
    def identity(self) -> float: ...               # ANI estimate in [0, 1]
    @property
    def reference_fraction(self) -> float: ...     # aligned fraction of the reference
    @property
    def query_fraction(self) -> float: ...         # aligned fraction of the query


class Sketch:
    """Read-only view of a sketched genome (sketch.rs); the arrays themselves stay on the device."""

    @property
    def name(self) -> str: ...
    @property
    def c(self) -> int: ...
    @property
    def amino_acid(self) -> bool: ...              # always False: pyskani exposes the DNA path only
