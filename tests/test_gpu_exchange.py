"""The one exchange step of the multi-GPU all-vs-all: the zero-copy exchange block on one GPU, and the
whole path over NCCL on two GPUs (skipped on a single-GPU box) against the single-GPU hit table."""
import os
import socket
import sys

import numpy as np
import pytest

from pyskani_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def small_genomes(n_fam=4, length=250_000, seed=3000):
    out = []
    for f in range(n_fam):
        base = synth.random_genome(length + 10_000 * f, seed + f)
        out.append([base.tobytes()])
        out.append([synth.mutate(base, 0.02, seed + 100 + f).tobytes()])
        out.append([synth.mutate(base, 0.07, seed + 200 + f).tobytes()])
        out.append([c.tobytes() for c in synth.fragment(synth.mutate(base, 0.04, seed + 300 + f), seed + 400 + f, lo=400, hi=30_000)])
    out.append([b"ACGT" * 60])             # below the contig gate: an empty sketch travels too
    return out


def test_exchange_block_pack_and_adopt():
    """skb_exchange_*: two 'ranks' (slices of one sketch list) pack their segments into one block; the sketches adopted
    from the block are identical to the originals, answer queries identically and keep the block alive on their own."""
    from pyskani_b200 import capi, parallel
    ctx = capi.Context(0)
    gs = ctx.sketch_batch(small_genomes())
    parts = [gs[:7], gs[7:]]
    sizes = [ctx.segment_size(p) for p in parts]
    assert all(s % 256 == 0 and m > 0 for s, m in sizes)
    for uniform_slack in (10.0, -1.0):                     # uniform stride forced, then exact back-to-back segments
        offs, total, stride = parallel.segment_layout([s for s, _ in sizes], uniform_slack)
        assert (stride != 0) == (uniform_slack > 0)
        ex = ctx.exchange(total)
        for o, p in zip(offs, parts):
            ex.pack(o, p)
        ctx.sync()
        got = ex.adopt(offs, [m for _, m in sizes], len(gs))
        ex.close()                                         # adopted sketches own the block from here on
        assert [len(x) for x in got] == [len(p) for p in parts]
        flat = got[0] + got[1]
        for a, b in zip(gs, flat):
            ia, ib = a.info(), b.info()
            assert (ia.n_seeds, ia.n_markers, ia.n_contigs, ia.total_len, ia.k, ia.c, ia.marker_c, ia.has_seeds) == \
                   (ib.n_seeds, ib.n_markers, ib.n_contigs, ib.total_len, ib.k, ib.c, ib.marker_c, ib.has_seeds)
            ea, eb = a.export(), b.export()
            for key in ea:
                assert np.array_equal(ea[key], eb[key]), key
        db_a, db_b = capi.Database(ctx), capi.Database(ctx)
        db_a.add_many(gs); db_b.add_many(flat)
        ha, na = db_a.query_array(gs)
        hb, nb = db_b.query_array(flat)
        assert na == nb and len(ha) >= 4 * 16 and np.array_equal(ha, hb)
        # mixed: adopted queries against the original database
        hc, _ = db_a.query_array(flat)
        assert np.array_equal(ha, hc)
    # error paths: misaligned offset, segment beyond the block, garbage where a descriptor should be
    ex = ctx.exchange(1 << 16)
    with pytest.raises(capi.SkbError):
        ex.pack(128, gs[:1])
    with pytest.raises(capi.SkbError):
        ex.pack(0, gs)                                     # does not fit
    with pytest.raises(capi.SkbError):
        ex.adopt([0], [64], 4)                             # nothing was packed there


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _nccl_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from pyskani_b200 import parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    be = parallel.CudaBackend(rank)
    genomes = small_genomes()
    out = []
    for _ in range(2):                                     # twice: the second exchange reuses slab memory of the first
        tm = {}
        table = parallel.all_vs_all(genomes, be, dist=dist, device=dev, timings=tm)
        out.append((table, tm))
    if rank == 0:
        q.put([(t, {k: v for k, v in tm.items() if not k.endswith("collective")}) for t, tm in out])
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_all_vs_all_equals_single_gpu():
    """The north-star split over NCCL: partition, sketch, exchange (pack + all-gather into the block + adopt), query own
    slice, gather hits.  The 2-GPU hit table must equal the single-GPU table bit for bit."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    from pyskani_b200 import parallel
    genomes = small_genomes()
    single = parallel.all_vs_all(genomes, parallel.CudaBackend(0))
    assert len(single) >= 4 * 16
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for table, tm in res:
        assert table.shape == single.shape and np.array_equal(table, single)
        assert tm["exchange_bytes_in"] > 0 and tm["exchange_allgather_ms"] > 0
