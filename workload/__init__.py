"""Synthetic genomes of BASELINE.json's configurations — neutral workload code.

Shared by both arms of bench.py (`--impl b200` and `--impl reference`), the tools and the parity gate, so that all of
them see byte-identical inputs.  Imports neither the product (pyskani_b200) nor the oracle.
ctypes front-end of workload/libsynthgen.so (synthgen.c); the calls release the GIL, so genomes are generated on a
thread pool.
"""
import ctypes as C
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

# BASELINE.json configs[2] / SURVEY.md §8d config 3: a family = one random root + mutants at these divergences
FAMILY_DIVERGENCES = (0.0, 0.01, 0.02, 0.03, 0.05, 0.07, 0.09, 0.11, 0.13, 0.15)
SEED0 = 0x5EED0000


def build(force=False):
    so = os.path.join(_HERE, "libsynthgen.so")
    src = os.path.join(_HERE, "synthgen.c")
    if force or not os.path.exists(so) or os.path.getmtime(src) > os.path.getmtime(so):
        env = dict(os.environ)
        env.pop("CC", None)
        subprocess.check_call(["make", "-C", _HERE, "-B", "libsynthgen.so"], stdout=subprocess.DEVNULL, env=env)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.wl_random_genome.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64]
        L.wl_random_genome.restype = None
        L.wl_mutate.argtypes = [C.c_void_p, C.c_uint64, C.c_double, C.c_uint64, C.c_void_p, C.c_uint64]
        L.wl_mutate.restype = C.c_uint64
        _LIB = L
    return _LIB


def slot_bytes(genome_len):
    """Capacity that holds any mutant of a genome_len genome (insertions are balanced by deletions; 2 % + 4 KB of slack),
    rounded to 16 so that consecutive slots keep the 16-byte alignment the device layout wants."""
    return (int(genome_len * 1.02) + 4096 + 15) // 16 * 16


def random_genome_into(dst_ptr, n, seed):
    lib().wl_random_genome(dst_ptr, n, seed)


def mutate_into(src_ptr, n, d, seed, dst_ptr, cap):
    return int(lib().wl_mutate(src_ptr, n, float(d), seed, dst_ptr, cap))


def random_genome(n, seed):
    out = np.empty(n, np.uint8)
    random_genome_into(out.ctypes.data, n, seed)
    return out


def mutate(genome, d, seed):
    genome = np.ascontiguousarray(genome, np.uint8)
    out = np.empty(slot_bytes(len(genome)), np.uint8)
    m = mutate_into(genome.ctypes.data, len(genome), d, seed, out.ctypes.data, len(out))
    return out[:m]


def genome_seed(family, member):
    return SEED0 + 1000 * family + member


def fill_families(buf, slot, genome_len, genome_ids, members=len(FAMILY_DIVERGENCES), divergences=FAMILY_DIVERGENCES,
                  threads=None):
    """Writes the genomes `genome_ids` (global id = family * members + member) into consecutive slots of `buf` (a uint8
    numpy array, e.g. a view of pinned memory): genome_ids[j] goes to buf[j * slot : ...].  Returns their lengths.
    Deterministic in the ids alone: every rank of a multi-process run can generate any subset and get the same bytes."""
    ids = list(genome_ids)
    lens = np.zeros(len(ids), np.uint64)
    by_family = {}
    for j, g in enumerate(ids):
        by_family.setdefault(g // members, []).append((j, g % members))
    base_ptr = buf.ctypes.data

    def one_family(item):
        fam, lst = item
        root = random_genome(genome_len, genome_seed(fam, 0))
        for j, m in lst:
            dst = base_ptr + j * slot
            if divergences[m] == 0.0:
                C.memmove(dst, root.ctypes.data, genome_len)
                lens[j] = genome_len
            else:
                lens[j] = mutate_into(root.ctypes.data, genome_len, divergences[m], genome_seed(fam, m), dst, slot - 16)

    with ThreadPoolExecutor(max_workers=threads or min(32, os.cpu_count() or 1)) as ex:
        list(ex.map(one_family, by_family.items()))
    return lens
