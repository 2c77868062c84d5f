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
                 k: int = 15, format: StorageFormat | None = None, model: PathLike | None = None) -> None:
        """``model`` (extension over pyskani): a gbdt-rs JSON dump of skani's learned-ANI regression; defaults to
        ``$PYSKANI_B200_MODEL``.  See ``query``."""

    # -- construction from disk (lib.rs:423-470): both load every sketch into device memory; ``open`` keeps the folder
    #    attached so that later ``sketch()`` calls append to it
    @classmethod
    def open(cls, path: PathLike, *, model: PathLike | None = None) -> Database: ...
    @classmethod
    def load(cls, path: PathLike, *, model: PathLike | None = None) -> Database: ...
    def set_model(self, path: PathLike | None) -> None:
        """Load (or, with None, drop) the learned-ANI regression: the serde_json dump of a gbdt-rs ``GBDT`` as skani embeds
        it (lib.rs:611-614; features: ANI %, std of window ANIs %, reference contig-length quantiles 90/50/10, query
        ones, aligned bases per chain, aligned bases).  It is evaluated on the device."""
    @property
    def has_model(self) -> bool: ...

    # -- the hot path
    def sketch(self, name: str, *contigs: Contig, seed: bool = True) -> None:
        """FracMinHash-sketch one genome on the GPU and add it (lib.rs:472-512). Contigs shorter than 500 bp are skipped."""
    def query(self, name: str, *contigs: Contig, seed: bool = True, learned_ani: bool | None = None,
              median: bool = False, robust: bool = False, cutoff: float | None = None,
              faster_small: bool = False) -> list[Hit]:
        """Sketch the query, screen it against every reference, chain the survivors, return hits with ANI > 0.1
        (lib.rs:549-660).

        ``learned_ani``: the reference resolves ``None`` to ``compression >= 70 and not median`` and then corrects the
        mean estimate with the regression model embedded in the skani crate (lib.rs:611-614).  Those weights are not
        part of pyskani's sources.  With a model loaded (``model=`` / ``set_model`` / ``$PYSKANI_B200_MODEL``) the
        correction runs on the device under the same rule; without one, ``None`` returns the UNCORRECTED estimate and
        emits a ``RuntimeWarning`` once per process (the reference's E. coli golden 0.9939 becomes 0.9946), and
        ``learned_ani=True`` raises ``RuntimeError``.  ``robust`` / ``median`` estimates are never corrected."""
    def sketch_many(self, items: _Seq[NamedGenome], *, seed: bool = True) -> None:
        """``sketch`` for many genomes in one device batch (no counterpart in the reference)."""
    def query_many(self, items: _Seq[NamedGenome], *, seed: bool = True, learned_ani: bool | None = None,
                   median: bool = False, robust: bool = False, cutoff: float | None = None,
                   faster_small: bool = False) -> list[list[Hit]]:
        """``query`` for many genomes at once; one hit list per item, in order."""

    # -- persistence (lib.rs:662-744)
    def save(self, path: PathLike, overwrite: bool = False, format: StorageFormat | None = None, *,
             strict_format: bool = False) -> None:
        """Writes ``markers.bin`` plus the sketches.  As in the reference (lib.rs:696-699) the two format names are
        swapped in this method: ``None`` / ``"consolidated"`` write one ``<name>.sketch`` per genome and ``"separated"``
        writes ``sketches.db`` + ``index.db``.  ``strict_format=True`` (extension) writes the layout the name says.
        ``Database.load`` / ``Database.open`` read either layout."""
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
