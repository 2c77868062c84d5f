"""N>1 host logic under gloo, world_size 2, on the CPU: partitioning, the sketch exchange and the hit gather of
pyskani_b200.parallel.  The device is replaced by a stand-in backend built on the CPU oracle (test infrastructure),
so what is verified here is the plumbing: a 2-rank run must return exactly the single-rank table."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleBackend:
    """Stand-in for CudaBackend: same interface, CPU oracle underneath."""

    def sketch(self, genomes, **params):
        import oracle
        return [(oracle.Sketch(g, **params), g) for g in genomes]

    def export(self, s):
        sk, _ = s
        kmer, pos, contig, canon = sk.seeds()
        return dict(kmer=kmer, pos=pos, contig=contig, canonical=canon, markers=sk.markers(), contig_lengths=sk.contig_lengths())

    def import_(self, e, **params):
        return ("imported", e)

    def query(self, db_sketches, query_sketches, **opts):
        import oracle
        # imported sketches carry only arrays; rebuild oracle sketches from the registry of raw genomes by identity of arrays
        refs = []
        for s in db_sketches:
            if s[0] == "imported":
                refs.append(self._rebuild(s[1]))
            else:
                refs.append(s[0])
        out = []
        for qi, (q, _) in enumerate(query_sketches):
            idx, res, _ = oracle.query(q, refs)
            out += [(qi, int(i), r.ani, r.af_query, r.af_ref) for i, r in zip(idx, res)]
        return out

    def _rebuild(self, e):
        # find the genome whose sketch exports to the same arrays
        import oracle
        for g in self.genomes:
            sk = oracle.Sketch(g)
            if sk.n_seeds == len(e["kmer"]) and np.array_equal(sk.seeds()[0], e["kmer"]) and np.array_equal(sk.markers(), e["markers"]):
                return sk
        raise AssertionError("exchanged sketch does not match any genome")


def make_genomes():
    from pyskani_b200 import synth
    out = []
    for f in range(3):
        base = synth.random_genome(120_000 + 20_000 * f, 500 + f)
        out.append([base.tobytes()])
        out.append([synth.mutate(base, 0.03, 600 + f).tobytes()])
        out.append([c.tobytes() for c in synth.fragment(synth.mutate(base, 0.06, 700 + f), 800 + f, lo=600, hi=30_000)])
    return out


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from pyskani_b200 import parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    genomes = make_genomes()
    be = OracleBackend()
    be.genomes = genomes
    table = parallel.all_vs_all(genomes, be, dist=dist, device="cpu")
    # the hit gather with uneven, empty and over-capacity row counts (one collective, repeated when a rank had more rows
    # than the remembered capacity)
    gathers = []
    for n0, n1 in ((5, 0), (0, 0), (3000, 7), (10, 2500)):
        n = n0 if rank == 0 else n1
        rows = np.arange(5 * n, dtype=np.float64).reshape(n, 5) + 1e6 * rank
        gathers.append(parallel.gather_hits(rows, dist, "cpu", sort=False))
    if rank == 0:
        q.put((table, gathers, dict(parallel._GATHER_ROWS)))
    dist.barrier()
    dist.destroy_process_group()


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_partition_is_balanced_and_deterministic():
    from pyskani_b200 import parallel
    sizes = [5, 9, 1, 7, 3, 3, 8, 2]
    parts = parallel.partition_by_size(sizes, 3)
    assert sorted(i for p in parts for i in p) == list(range(8))
    loads = [sum(sizes[i] for i in p) for p in parts]
    assert max(loads) - min(loads) <= max(sizes)
    assert parts == parallel.partition_by_size(sizes, 3)
    assert parallel.partition_by_size([], 2) == [[], []]
    assert parallel.partition_by_size([4], 4) == [[0], [], [], []]


def test_segment_layout():
    """exchange block layout: a uniform stride (one in-place all-gather) when the segments are nearly equal, exact
    back-to-back segments (grouped broadcasts) when padding to the largest would waste more than 10 %"""
    from pyskani_b200 import parallel
    offs, total, stride = parallel.segment_layout([2560, 2816, 2560, 2816])
    assert stride == 2816 and offs == [0, 2816, 5632, 8448] and total == 11264
    offs, total, stride = parallel.segment_layout([256, 4096, 256])
    assert stride == 0 and offs == [0, 256, 4352] and total == 4608
    offs, total, stride = parallel.segment_layout([0, 0])
    assert total >= 256 and len(offs) == 2
    offs, end, stride = parallel.segment_layout([512, 512], base=4096)        # a second region behind the first
    assert offs == [4096, 4608] and end == 5120 and stride == 512
    for segs in ([512] * 8, [768, 256, 1024], [256]):
        offs, total, stride = parallel.segment_layout(segs)
        for r, (o, s) in enumerate(zip(offs, segs)):
            assert o % 256 == 0 and o + s <= total
            assert all(o + s <= offs[j] for j in range(r + 1, len(segs)))      # segments never overlap


def test_pack_roundtrip():
    from pyskani_b200 import parallel
    rng = np.random.default_rng(0)
    sk = []
    for n in (0, 5, 1000):
        sk.append(dict(kmer=rng.integers(0, 2**30, n).astype(np.uint64), pos=rng.integers(0, 10**6, n).astype(np.uint32),
                       contig=rng.integers(0, 9, n).astype(np.uint32), canonical=rng.integers(0, 2, n).astype(np.uint8),
                       markers=np.sort(rng.integers(0, 2**42, n // 8).astype(np.uint64)), contig_lengths=np.array([n, 7], np.uint32)))
    h, p = parallel.pack_sketches(sk)
    back = parallel.unpack_sketches(h, p)
    assert len(back) == 3
    for a, b in zip(sk, back):
        for k in a:
            assert np.array_equal(a[k], b[k]) and a[k].dtype == b[k].dtype


def test_two_ranks_equal_one_rank():
    import torch.multiprocessing as mp
    from pyskani_b200 import parallel
    genomes = make_genomes()
    be = OracleBackend()
    be.genomes = genomes
    single = parallel.all_vs_all(genomes, be)
    assert len(single) >= 3 * 9 - 2     # every intra-family ordered pair (incl. self) is a hit at these divergences
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    table, gathers, caps = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for (n0, n1), g in zip(((5, 0), (0, 0), (3000, 7), (10, 2500)), gathers):
        want = np.concatenate([np.arange(5 * n0, dtype=np.float64).reshape(n0, 5), np.arange(5 * n1, dtype=np.float64).reshape(n1, 5) + 1e6])
        assert g.shape == want.shape and np.array_equal(g, want)
    assert caps[2] >= 3000
    assert table.shape == single.shape
    assert np.array_equal(table[:, :2], single[:, :2])
    assert np.allclose(table[:, 2:], single[:, 2:], rtol=0, atol=0)
