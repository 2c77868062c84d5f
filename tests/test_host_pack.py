"""pyskani_b200/csrc/host_pack.cpp on the CPU: the 2-bit words the ingest pipeline of skb_sketch_batch produces on the host
must be the words kmer_bits.cuh::pack16 produces on the device, for every byte value, alignment and length."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("shim") / "host_pack_shim.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-pthread", "-o", so,
                           os.path.join(HERE, "host_shim", "host_pack_shim.cpp")])
    L = C.CDLL(so)
    L.shim_host_pack.restype = C.c_uint64
    L.shim_host_pack.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_int]
    L.shim_host_pack_isa.restype = C.c_char_p
    L.shim_device_pack.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
    L.shim_team_pack.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint, C.c_uint64, C.c_int]
    L.shim_cpu_count.restype = C.c_uint
    return L


ISAS = (0, 1, 2, 3, 4)    # dispatching entry, scalar, AVX2, AVX-512 VBMI, dispatching entry with ordinary stores


def messy(n, seed):
    """mostly ACGT in both cases, with runs of N, IUPAC codes, newlines and arbitrary bytes"""
    rng = np.random.default_rng(seed)
    s = np.frombuffer(b"ACGTacgt", np.uint8)[rng.integers(0, 8, n)].copy()
    k = max(1, n // 50)
    s[rng.integers(0, n, k)] = np.frombuffer(b"NnRYKMSWBDHV\n-.*", np.uint8)[rng.integers(0, 16, k)]
    s[rng.integers(0, n, k)] = rng.integers(0, 256, k).astype(np.uint8)
    if n > 4000:
        s[1000:1700] = ord("N")
    return s


def host_words(L, buf, off, n, isa):
    out = np.full((n + 15) // 16 + 9, 0xDEADBEEF, np.uint32)        # written at an odd word offset: unaligned stores
    if L.shim_host_pack(buf.ctypes.data + off, n, out.ctypes.data + 4, isa) == 0:
        return None                                                 # instruction set not available here
    assert out[0] == 0xDEADBEEF and (out[1 + (n + 15) // 16:] == 0xDEADBEEF).all(), "wrote outside its range"
    return out[1:1 + (n + 15) // 16]


def device_words(L, buf, off, n):
    out = np.zeros((n + 15) // 16, np.uint32)
    L.shim_device_pack(buf.ctypes.data + off, n, out.ctypes.data)
    return out


def test_every_byte_value(shim):
    for b in range(256):
        buf = np.full(300, ord("G"), np.uint8)
        buf[::7] = b
        for isa in ISAS:
            got = host_words(shim, buf, 0, 300, isa)
            assert got is None or (got == device_words(shim, buf, 0, 300)).all(), (b, isa)


@pytest.mark.parametrize("n", [0, 1, 15, 16, 17, 127, 128, 129, 255, 1000, 4096 + 5, 100_003])
def test_lengths_and_alignments(shim, n):
    buf = messy(n + 64, n)
    for off in (0, 1, 3, 16, 31):
        want = device_words(shim, buf, off, n)
        for isa in ISAS:
            got = host_words(shim, buf, off, n, isa)
            assert got is None or (got == want).all(), (n, off, isa)


def test_team_rounds(shim):
    n = 3_000_000 + 7
    buf = messy(n, 5)
    want = device_words(shim, buf, 0, n)
    for threads in (1, 3, 8):
        out = np.zeros((n + 15) // 16, np.uint32)
        assert shim.shim_team_pack(buf.ctypes.data, n, out.ctypes.data, threads, 1 << 16, 5) == 5
        assert (out == want).all()
    assert shim.shim_cpu_count() >= 1
    assert shim.shim_host_pack_isa() in (b"avx512vbmi", b"avx2", b"scalar")
