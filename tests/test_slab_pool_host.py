"""pyskani_b200/csrc/slab_pool.h (the allocator behind every sketch array) compiled for the host with malloc as the raw
allocator: bump allocation, roll-back of the most recent allocation, reuse of idle slabs, growth policy, failure."""
import ctypes as C
import os
import random
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
MB = 1 << 20


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("shim") / "slab_pool_shim.so")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-o", so,
                           os.path.join(HERE, "host_shim", "slab_pool_shim.cpp")])
    L = C.CDLL(so)
    L.pool_new.restype = C.c_void_p
    L.pool_delete.argtypes = [C.c_void_p]
    L.pool_fail_above.argtypes = [C.c_long]
    L.pool_raw_calls.restype = C.c_long
    L.pool_reserved.restype = C.c_ulonglong; L.pool_reserved.argtypes = [C.c_void_p]
    L.pool_slabs.restype = C.c_ulonglong; L.pool_slabs.argtypes = [C.c_void_p]
    L.pool_alloc.restype = C.c_void_p
    L.pool_alloc.argtypes = [C.c_void_p, C.c_ulonglong, C.POINTER(C.c_ulonglong), C.POINTER(C.c_long)]
    L.pool_free.argtypes = [C.c_void_p, C.c_void_p]
    for f in ("slab_used", "slab_cap"):
        getattr(L, f).restype = C.c_ulonglong; getattr(L, f).argtypes = [C.c_void_p, C.c_long]
    L.slab_live.restype = C.c_uint; L.slab_live.argtypes = [C.c_void_p, C.c_long]
    return L


class Pool:
    def __init__(self, L):
        self.L, self.h = L, L.pool_new()

    def alloc(self, n):
        addr, slab = C.c_ulonglong(), C.c_long()
        hd = self.L.pool_alloc(self.h, n, C.byref(addr), C.byref(slab))
        return None if not hd else (hd, addr.value, slab.value, n)

    def free(self, a):
        self.L.pool_free(self.h, a[0])

    def close(self):
        self.L.pool_delete(self.h)


def test_bump_alignment_and_rollback(lib):
    p = Pool(lib)
    a = p.alloc(1000); b = p.alloc(1); c = p.alloc(513)
    assert a[2] == b[2] == c[2] == 0 and lib.pool_slabs(p.h) == 1 and lib.slab_cap(p.h, 0) == 64 * MB
    assert a[1] % 512 == 0 and b[1] == a[1] + 1024 and c[1] == b[1] + 512          # rounded to the 512-byte granule
    assert lib.slab_used(p.h, 0) == 1024 + 512 + 1024 and lib.slab_live(p.h, 0) == 3
    p.free(c)                                                                      # most recent: the bump pointer rolls back
    assert lib.slab_used(p.h, 0) == 1024 + 512
    d = p.alloc(100)
    assert d[1] == c[1]
    p.free(a)                                                                      # not the most recent: space stays taken
    assert lib.slab_used(p.h, 0) == 1024 + 512 + 512 and lib.slab_live(p.h, 0) == 2
    p.free(d); p.free(b)                                                           # idle slab restarts from zero
    assert lib.slab_used(p.h, 0) == 0 and lib.slab_live(p.h, 0) == 0
    e = p.alloc(10)
    assert e[1] == a[1] and lib.pool_raw_calls() == 1
    p.free(e); p.close()


def test_growth_reuse_and_query_sketch_pattern(lib):
    p = Pool(lib)
    db = [p.alloc(40 * MB) for _ in range(5)]           # the 64 MB slab holds one, the 128 MB slab three, then 256 MB
    caps = [lib.slab_cap(p.h, i) for i in range(lib.pool_slabs(p.h))]
    assert caps == [64 * MB, 128 * MB, 256 * MB] and [x[2] for x in db] == [0, 1, 1, 1, 2]
    calls = lib.pool_raw_calls()
    for _ in range(1000):                                # Database.query: a temporary sketch per call, freed right after
        q = p.alloc(1 * MB)
        p.free(q)
    assert lib.pool_raw_calls() == calls and lib.slab_used(p.h, 2) == 40 * MB
    big = p.alloc(700 * MB)                              # larger than the next slab size: gets a slab of its own size
    assert lib.slab_cap(p.h, big[2]) == 700 * MB
    for x in db:
        p.free(x)
    # the idle slabs are reused, smallest fitting first, instead of asking the raw allocator again
    calls = lib.pool_raw_calls()
    y = p.alloc(100 * MB)
    assert lib.pool_raw_calls() == calls and lib.slab_cap(p.h, y[2]) == 128 * MB
    p.free(y); p.free(big); p.close()


def test_raw_allocator_failure(lib):
    p = Pool(lib)
    lib.pool_fail_above(10 * MB)
    assert p.alloc(11 * MB) is None                      # neither the 64 MB slab nor the exact size can be had
    a = p.alloc(5 * MB)                                  # the 64 MB slab fails, the exact-size retry succeeds
    assert a is not None and lib.slab_cap(p.h, a[2]) == 5 * MB
    p.free(a); p.close()


def test_random_traffic_never_overlaps(lib):
    rng = random.Random(5)
    p = Pool(lib)
    live = []
    for step in range(4000):
        if live and (rng.random() < 0.45 or len(live) > 60):
            p.free(live.pop(rng.randrange(len(live)) if rng.random() < 0.5 else -1))
        else:
            a = p.alloc(rng.choice([1, 700, 4096, 100_000, 3 * MB, 20 * MB]))
            assert a is not None
            live.append(a)
        if step % 200 == 0:
            spans = sorted((a[1], a[1] + ((a[3] + 511) // 512) * 512) for a in live)
            assert all(x[1] <= y[0] for x, y in zip(spans, spans[1:]))
    for a in live:
        p.free(a)
    assert all(lib.slab_live(p.h, i) == 0 and lib.slab_used(p.h, i) == 0 for i in range(lib.pool_slabs(p.h)))
    p.close()
