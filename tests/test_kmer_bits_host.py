"""pyskani_b200/csrc/kmer_bits.cuh compiled for the host and checked against the oracle.

The seeding kernel's bit manipulation (ASCII->2-bit packing, funnel-shift k-mer extraction, reverse
complement by bit reversal, the hash) is __host__ __device__ code; this test runs the very same source
on the CPU so that it is verified on every `-m "not gpu"` run.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("shim") / "kmer_bits_shim.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so,
                           os.path.join(HERE, "host_shim", "kmer_bits_shim.cpp")])
    L = C.CDLL(so)
    L.shim_hash.restype = C.c_uint64
    L.shim_hash.argtypes = [C.c_uint64]
    L.shim_pack16.restype = C.c_uint32
    L.shim_pack16.argtypes = [C.c_char_p]
    L.shim_scan.restype = C.c_int64
    L.shim_scan.argtypes = [C.c_char_p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                            C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(C.c_int64)]
    return L


def scan(L, seq, k=15, c=125, marker_c=1000):
    cap = len(seq)
    kmer = np.empty(cap, np.uint32); pos = np.empty(cap, np.uint32); canon = np.empty(cap, np.uint8)
    markers = np.empty(cap, np.uint64); nm = C.c_int64(0)
    ns = L.shim_scan(seq, len(seq), k, c, marker_c, kmer.ctypes.data, pos.ctypes.data, canon.ctypes.data,
                     cap, markers.ctypes.data, C.byref(nm))
    return kmer[:ns], pos[:ns], canon[:ns], np.unique(markers[:nm.value])


def rand_seq(n, seed, alphabet=b"ACGT"):
    rng = np.random.default_rng(seed)
    return np.frombuffer(alphabet, np.uint8)[rng.integers(0, len(alphabet), n)].tobytes()


def test_hash_matches_oracle(shim):
    rng = np.random.default_rng(0)
    for x in [0, 1, 2**30 - 1, 2**42 - 1] + [int(v) for v in rng.integers(0, 2**42, 200)]:
        assert shim.shim_hash(x) == oracle.mm_hash64(x)


def test_pack16_all_bytes(shim):
    code = {ord("A"): 0, ord("C"): 1, ord("G"): 2, ord("T"): 3, ord("a"): 0, ord("c"): 1, ord("g"): 2, ord("t"): 3}
    for b in range(256):
        for slot in (0, 5, 15):
            buf = bytearray(b"A" * 16)
            buf[slot] = b
            w = shim.shim_pack16(bytes(buf))
            assert w == code.get(b, 0) << (30 - 2 * slot), (b, slot, hex(w))


@pytest.mark.parametrize("n,seed,k,c,mc", [(600, 1, 15, 125, 1000), (100_003, 2, 15, 125, 1000),
                                            (50_000, 3, 13, 30, 200), (50_001, 4, 16, 50, 500),
                                            (20_000, 5, 15, 1, 1)])
def test_scan_matches_oracle(shim, n, seed, k, c, mc):
    s = rand_seq(n, seed)
    kmer, pos, canon, markers = scan(shim, s, k, c, mc)
    O = oracle.Sketch([s], k=k, c=c, marker_c=mc)
    ok, op, _, oc = O.seeds()
    o = np.lexsort((pos, kmer))
    assert np.array_equal(kmer[o].astype(np.uint64), ok)
    assert np.array_equal(pos[o], op) and np.array_equal(canon[o], oc)
    assert np.array_equal(markers, O.markers())


def test_scan_with_junk_bytes(shim):
    s = rand_seq(40_000, 9, b"ACGTNacgtnRYKM-*")
    kmer, pos, canon, markers = scan(shim, s)
    O = oracle.Sketch([s])
    ok, op, _, oc = O.seeds()
    o = np.lexsort((pos, kmer))
    assert np.array_equal(kmer[o].astype(np.uint64), ok) and np.array_equal(pos[o], op)
    assert np.array_equal(canon[o], oc) and np.array_equal(markers, O.markers())


def test_scan_ecoli_slice(shim, ecoli):
    s = ecoli[1][2_000_000:2_400_000]
    kmer, pos, canon, markers = scan(shim, s)
    O = oracle.Sketch([s])
    ok, op, _, oc = O.seeds()
    o = np.lexsort((pos, kmer))
    assert np.array_equal(kmer[o].astype(np.uint64), ok) and np.array_equal(pos[o], op)
    assert np.array_equal(markers, O.markers())


def test_carry_form_of_the_high_word_comparison():
    """seed_kernels.cu::push_miss: with M = 2^32 - 1 the carry of hi(h * M) + (2^32 - thr) is [h > thr] for every
    threshold >= 1 (the form that lets one IMAD.HI produce the comparison as its carry predicate), and shifting the carries
    in from the last position to the first leaves bit e = position e missed."""
    rng = np.random.default_rng(5)
    M = (1 << 32) - 1
    thrs = [1, 2, 34359738, (1 << 31) - 1, 1 << 31, (1 << 32) - 2, (1 << 32) - 1] + [int(t) for t in rng.integers(1, 1 << 32, 200)]
    for thr in thrs:
        comp = (-thr) & M
        hs = [0, 1, thr - 1, thr, min(thr + 1, M), M] + [int(h) for h in rng.integers(0, 1 << 32, 300)]
        for h in hs:
            carry = (((h * M) >> 32) + comp) >> 32
            assert carry == (1 if h > thr else 0), (thr, h)
    thr, comp = 34359738, (-34359738) & M
    hs = [int(h) for h in rng.integers(0, 1 << 27, 16)]               # one word: positions 0..15
    miss = 0
    for e in range(15, -1, -1):
        miss = (miss * 2 + ((((hs[e] * M) >> 32) + comp) >> 32)) & M
    assert (~miss) & 0xFFFF == sum(1 << e for e in range(16) if hs[e] <= thr)
