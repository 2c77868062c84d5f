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
    """skb_exchange_*: two 'ranks' (slices of one sketch list) pack their head and body segments into one block; the
    sketches adopted from the block are identical to the originals, answer queries identically and keep the block alive on
    their own.  Full transfer and reference-only transfer (no query-side arrays: half of the bytes)."""
    from pyskani_b200 import capi, parallel
    ctx = capi.Context(0)
    gs = ctx.sketch_batch(small_genomes())
    parts = [gs[:7], gs[7:]]
    db_a = capi.Database(ctx)
    db_a.add_many(gs)
    ha, na = db_a.query_array(gs)
    assert len(ha) >= 4 * 16
    full_bytes = None
    for ref_only in (False, True):
        sizes = [ctx.segment_size(p, ref_only) for p in parts]
        assert all(h % 256 == 0 and b % 256 == 0 and m > 0 for h, b, m in sizes)
        if ref_only:
            assert sum(b for _, b, _ in sizes) < 0.6 * full_bytes          # bodies without the position-order arrays
        else:
            full_bytes = sum(b for _, b, _ in sizes)
        for uniform_slack in (10.0, -1.0):                     # uniform stride forced, then exact back-to-back segments
            h_offs, h_end, h_stride = parallel.segment_layout([h for h, _, _ in sizes], uniform_slack)
            b_offs, total, b_stride = parallel.segment_layout([b for _, b, _ in sizes], uniform_slack, base=h_end)
            assert (h_stride != 0) == (uniform_slack > 0) == (b_stride != 0)
            ex = ctx.exchange(total)
            for ho, bo, p in zip(h_offs, b_offs, parts):
                ex.pack(ho, bo, p, ref_only)
            # the bodies "arrive" on the context's own stream here; order_after records the event that chaining waits for
            ex.order_after(ctx.stream, bodies=False)
            ex.order_after(ctx.stream, bodies=True)
            got = ex.adopt(h_offs, b_offs, [m for _, _, m in sizes], len(gs))
            ex.close()                                         # adopted sketches own the block from here on
            assert [len(x) for x in got] == [len(p) for p in parts]
            flat = got[0] + got[1]
            for a, b in zip(gs, flat):
                ia, ib = a.info(), b.info()
                assert (ia.n_seeds, ia.n_markers, ia.n_contigs, ia.total_len, ia.k, ia.c, ia.marker_c, ia.has_seeds) == \
                       (ib.n_seeds, ib.n_markers, ib.n_contigs, ib.total_len, ib.k, ib.c, ib.marker_c, ib.has_seeds)
                assert ib.reference_only == int(ref_only) and ia.reference_only == 0
                ea, eb = a.export(), b.export()
                for key in ea:
                    assert np.array_equal(ea[key], eb[key]), key
            db_b = capi.Database(ctx)
            db_b.add_many(flat)
            hb, nb = db_b.query_array(gs)                      # originals as queries against the adopted database
            assert na == nb and np.array_equal(ha, hb)
            ok_a, sh_a = db_a.screen(gs[:3])
            ok_b, sh_b = db_b.screen(gs[:3])
            assert np.array_equal(ok_a, ok_b) and np.array_equal(sh_a, sh_b)
            if ref_only:
                with pytest.raises(capi.SkbError) as e:        # reference-only sketches cannot be queries
                    db_a.query_array(flat[:1])
                assert e.value.code == capi.SKB_ERR_ARG
                # ... but they can travel on: re-packed (still reference-only) and adopted again
                h2, b2, m2 = ctx.segment_size(flat[:3], False)
                ex2 = ctx.exchange(h2 + b2)
                ex2.pack(0, h2, flat[:3], False)
                (again,) = ex2.adopt([0], [h2], [m2], 3)
                assert all(x.info().reference_only == 1 for x in again)
                assert np.array_equal(again[1].export()["kmer"], gs[1].export()["kmer"])
            else:
                hc, _ = db_a.query_array(flat)                 # adopted sketches as queries
                assert np.array_equal(ha, hc)
    # error paths: misaligned offset, segment beyond the block, garbage where a descriptor should be
    ex = ctx.exchange(1 << 16)
    with pytest.raises(capi.SkbError):
        ex.pack(128, 4096, gs[:1])
    with pytest.raises(capi.SkbError):
        ex.pack(0, 4096, gs)                                   # does not fit
    with pytest.raises(capi.SkbError):
        ex.adopt([0], [4096], [64], 4)                         # nothing was packed there


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


def test_pipelined_all_vs_all_equals_the_plain_one():
    """parallel.all_vs_all_pipelined: sketch batches arrive one after the other on one context while a worker thread queries
    them on a SECOND context of the same device (new batch x database so far, both directions).  The table must equal the
    plain all-vs-all of the same genomes, for several batch splits (incl. an empty batch and a single one), and sketches of
    one context must be usable in a database of the other."""
    from pyskani_b200 import capi, parallel
    ctx_s, ctx_q = capi.Context(0), capi.Context(0)
    genomes = small_genomes(n_fam=5)
    gs = ctx_s.sketch_batch(genomes)
    db = capi.Database(ctx_s)
    db.add_many(gs)
    h, _ = db.query_array(gs)
    want = np.zeros((len(h), 5))
    want[:, 0] = h["query_index"]; want[:, 1] = h["ref_index"]; want[:, 2] = h["ani"]; want[:, 3] = h["af_query"]; want[:, 4] = h["af_ref"]
    want = parallel.sort_hits(want)
    assert len(want) >= 5 * 16
    bq = parallel.CudaBackend(0, ctx=ctx_q)
    for cuts in ([0, 7, 12, len(genomes)], [0, len(genomes)], [0, 1, 1, 9, len(genomes)], [0, 3, 6, 9, 12, 15, 18, len(genomes)]):
        def batches():
            for a, b in zip(cuts[:-1], cuts[1:]):
                yield ctx_s.sketch_batch(genomes[a:b])          # sketched while the worker queries the earlier ones
        tm = {}
        got = parallel.all_vs_all_pipelined(batches(), bq, timings=tm)
        assert got.shape == want.shape and np.array_equal(got, want), cuts
        assert tm["local_hits"] == len(want) and tm["screened_in"] >= len(want)
    # an error on the worker thread reaches the caller
    other = ctx_s.sketch_batch(genomes[:2], c=30)
    with pytest.raises(capi.SkbError):
        parallel.all_vs_all_pipelined(iter([gs[:3], other]), bq)


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
