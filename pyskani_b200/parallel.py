"""Multi-GPU all-vs-all (BASELINE.json configs 3 and 5).

One process per GPU (torch.distributed: NCCL on GPUs, gloo in the CPU tests).  The pair matrix partitions by
query, so the only exchange step is the sketch database:

  1. genomes are split across ranks, balanced by total bases        (partition_by_size)
  2. every rank sketches its share on its own GPU                    (backend.sketch)
  3. sketches are exchanged ONCE, device to device, without staging  (exchange_sketches_device):
       a. a 3-number all-gather tells every rank the segment size of every other rank
       b. every rank allocates ONE block of sketch storage with a segment per rank and packs its own sketches into its
          segment (skb_exchange_pack: one gather kernel)
       c. ONE collective fills the other segments in place: ncclAllGather straight into the block when the segments
          are (nearly) the same size, a group of per-rank broadcasts on exact sizes otherwise
       d. the peers' sketches are adopted as views into the block (skb_exchange_adopt): no unpack copy
     Fallback / CPU tests: an all-gather of the exported host SoA    (exchange_sketches)
  4. every rank builds the full database and queries ITS genomes against it
  5. hits are gathered on rank 0                                     (gather_hits)

There is no collective inside screen / chain / ANI.  `backend` is the object that talks to the device:
CudaBackend (libskb through pyskani_b200.capi) in production; the tests plug a CPU stand-in to exercise the
partitioning / exchange / gather logic under gloo.
"""
import time

import numpy as np

_FIELDS = (("kmer", np.uint64), ("pos", np.uint32), ("contig", np.uint32), ("canonical", np.uint8),
           ("markers", np.uint64), ("contig_lengths", np.uint32))


def partition_by_size(sizes, world):
    """Greedy longest-processing-time split of item indices over `world` ranks; returns a list of index lists.
    Deterministic (ties broken by index) so that every rank computes the same plan without communicating."""
    order = sorted(range(len(sizes)), key=lambda i: (-int(sizes[i]), i))
    loads = [0] * world
    parts = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda x: (loads[x], x))
        parts[r].append(i)
        loads[r] += int(sizes[i])
    return [sorted(p) for p in parts]


def pack_sketches(sketches):
    """list of export() dicts -> (header int64 array, one flat uint8 payload)."""
    header = [len(sketches)]
    chunks = []
    for e in sketches:
        for name, dt in _FIELDS:
            a = np.ascontiguousarray(e[name], dt)
            header.append(a.size)
            chunks.append(a.view(np.uint8).reshape(-1))
    payload = np.concatenate(chunks) if chunks else np.zeros(0, np.uint8)
    return np.asarray(header, np.int64), payload


def unpack_sketches(header, payload):
    n = int(header[0])
    out, h, off = [], 1, 0
    for _ in range(n):
        e = {}
        for name, dt in _FIELDS:
            cnt = int(header[h]); h += 1
            nbytes = cnt * np.dtype(dt).itemsize
            e[name] = payload[off:off + nbytes].view(dt).copy()
            off += nbytes
        out.append(e)
    return out


def _all_gather_var(arr, dist, device):
    """all-gather of 1-D numpy arrays of different lengths (pad to the max, trim after)."""
    import torch
    world = dist.get_world_size()
    t = torch.from_numpy(np.ascontiguousarray(arr)).to(device)
    n = torch.tensor([t.numel()], dtype=torch.int64, device=device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    m = max(sizes + [1])
    pad = torch.zeros(m, dtype=t.dtype, device=device)
    pad[:t.numel()] = t
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    return [b[:s].cpu().numpy() for b, s in zip(bufs, sizes)]


def exchange_sketches(local_exports, dist, device="cpu"):
    """Host-staged exchange (fallback and gloo tests): every rank contributes its exported sketches, every rank receives
    all of them (list over ranks of lists of export dicts)."""
    header, payload = pack_sketches(local_exports)
    headers = _all_gather_var(header, dist, device)
    payloads = _all_gather_var(payload, dist, device)
    return [unpack_sketches(h, p) for h, p in zip(headers, payloads)]


class _DevBlock:
    """A raw device address range as something torch.as_tensor can wrap without copying."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def segment_layout(seg_bytes, uniform_slack=0.10):
    """Offsets of the per-rank segments inside the exchange block.  Segments of (nearly) equal size are laid out at a
    uniform stride so that ONE in-place ncclAllGather can fill the block; otherwise they are packed back to back and
    a group of exact-size broadcasts is used.  Returns (offsets, total bytes, uniform stride or 0)."""
    world = len(seg_bytes)
    stride = max(seg_bytes + [256])
    if stride * world <= (1.0 + uniform_slack) * max(sum(seg_bytes), 1):
        return [r * stride for r in range(world)], stride * world, stride
    offs, cur = [], 0
    for s in seg_bytes:
        offs.append(cur)
        cur += s
    return offs, max(cur, 256), 0


def exchange_sketches_device(backend, local_sketches, dist, device, timings=None):
    """The one data-path collective, device resident: returns a list over ranks of lists of sketch handles living on this
    rank's GPU (this rank's entry is `local_sketches` itself; the others are views into one exchange block).
    `timings` (dict, optional) receives pack / all-gather / adopt milliseconds and the byte counts."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    ctx = backend.ctx
    t0 = time.perf_counter()
    seg, mb = ctx.segment_size(local_sketches)
    mine = torch.tensor([seg, mb, len(local_sketches)], dtype=torch.int64, device=device)
    allsz = torch.empty(3 * world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(allsz, mine)
    sz = allsz.tolist()
    segs, metas, counts = sz[0::3], sz[1::3], sz[2::3]
    offs, total, stride = segment_layout(segs)
    ex = ctx.exchange(total)
    t1 = time.perf_counter()
    ex.pack(offs[rank], local_sketches)                       # asynchronous on libskb's stream
    lib_stream = torch.cuda.ExternalStream(ctx.stream, device=device)
    cur = torch.cuda.current_stream(device)
    ev_in, ev_a, ev_b = torch.cuda.Event(), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev_in.record(lib_stream)
    cur.wait_event(ev_in)                                     # the collective starts when this rank's segment is packed
    block = torch.as_tensor(_DevBlock(ex.ptr, total), device=device)
    ev_a.record(cur)
    if stride:
        dist.all_gather_into_tensor(block, block[offs[rank]:offs[rank] + stride])          # in place
    else:
        views = [block[offs[r]:offs[r] + segs[r]] for r in range(world)]
        dist.all_gather(views, views[rank])                                                # exact sizes: grouped broadcasts
    ev_b.record(cur)
    lib_stream.wait_event(ev_b)                               # libskb reads the block only after the collective
    t2 = time.perf_counter()
    peers = [r for r in range(world) if r != rank]
    adopted = ex.adopt([offs[r] for r in peers], [metas[r] for r in peers], sum(counts[r] for r in peers))
    t3 = time.perf_counter()
    ex.close()                                                # the block now belongs to the adopted sketches
    out = [None] * world
    out[rank] = local_sketches
    for r, lst in zip(peers, adopted):
        out[r] = lst
    if timings is not None:
        timings["exchange_sizes_ms"] = 1e3 * (t1 - t0)
        timings["exchange_pack_enqueue_ms"] = 1e3 * (t2 - t1)
        timings["exchange_wait_adopt_ms"] = 1e3 * (t3 - t2)
        timings["exchange_allgather_ms"] = ev_a.elapsed_time(ev_b)        # adopt() synchronised the stream behind ev_b
        timings["exchange_ms"] = 1e3 * (t3 - t0)
        timings["exchange_bytes_in"] = int(sum(segs) - segs[rank])
        timings["exchange_bytes_total"] = int(sum(segs))
        timings["exchange_collective"] = "ncclAllGather in place (uniform stride)" if stride else "grouped ncclBroadcast on exact sizes"
    return out


def gather_hits(local_hits, dist, device="cpu"):
    """hits: (n, 5) float64 rows [query_global, ref_global, ani, af_query, af_ref]; returned on rank 0 sorted by (query, ref)."""
    flat = np.asarray(local_hits, np.float64).reshape(-1)
    parts = _all_gather_var(flat, dist, device)
    allh = np.concatenate(parts).reshape(-1, 5) if parts else np.zeros((0, 5))
    order = np.lexsort((allh[:, 1], allh[:, 0]))
    return allh[order]


class CudaBackend:
    """libskb on this rank's GPU."""

    def __init__(self, device_index, ctx=None):
        from . import capi
        self.capi = capi
        self.ctx = ctx if ctx is not None else capi.Context(device_index)

    def sketch(self, genomes, **params):
        return self.ctx.sketch_batch(genomes, **params)

    def export(self, sketch):
        return sketch.export()

    def import_(self, e, **params):
        return self.ctx.import_sketch(e["kmer"], e["pos"], e["contig"], e["canonical"], e["markers"], e["contig_lengths"], **params)

    def pack(self, sketches, device):
        import torch
        pb, _ = self.ctx.pack_size(sketches)
        payload = torch.empty(max(pb, 16), dtype=torch.uint8, device=device)
        torch.cuda.synchronize(device)
        meta = self.ctx.pack(sketches, payload.data_ptr(), payload.numel())
        return payload[:max(pb, 0)] if pb else payload[:0], meta

    def unpack(self, meta, payload):
        return self.ctx.unpack(meta, payload.data_ptr() if payload.numel() else None, payload.numel())

    def query(self, db_sketches, query_sketches, max_pairs_per_call=1 << 27, stats=None, **opts):
        """Rows (query index, ref index, ani, af_query, af_ref) as a float64 array; the queries go through skb_db_query
        in slices of at most max_pairs_per_call pairs."""
        db = self.capi.Database(self.ctx)
        db.add_many(list(db_sketches))
        nr = max(1, len(db_sketches))
        step = max(1, max_pairs_per_call // nr)
        rows, n_in, screen_ms, chain_ms = [], 0, 0.0, 0.0
        for q0 in range(0, len(query_sketches), step):
            h, k = db.query_array(query_sketches[q0:q0 + step], **opts)
            n_in += k
            st = self.ctx.stats()
            screen_ms += st.screen_ms; chain_ms += st.chain_ms
            if len(h):
                a = np.empty((len(h), 5), np.float64)
                a[:, 0] = h["query_index"] + q0; a[:, 1] = h["ref_index"]
                a[:, 2] = h["ani"]; a[:, 3] = h["af_query"]; a[:, 4] = h["af_ref"]
                rows.append(a)
        if stats is not None:
            stats["screened_in"] = stats.get("screened_in", 0) + n_in
            stats["screen_ms"] = stats.get("screen_ms", 0.0) + screen_ms
            stats["chain_ms"] = stats.get("chain_ms", 0.0) + chain_ms
        return np.concatenate(rows) if rows else np.zeros((0, 5))


def query_and_gather(backend, local_sketches, mine, plan, dist=None, device="cpu", query_opts=None, timings=None,
                     import_params=None):
    """Steps 3-5 for sketches that already exist on this rank: exchange, query this rank's genomes against the full
    database, gather the hit table [query, ref, ani, af_query, af_ref] (global ids, sorted) on rank 0."""
    query_opts = query_opts or {}
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    n_total = sum(len(p) for p in plan)
    full = [None] * n_total
    on_gpu = world > 1 and hasattr(backend, "ctx") and str(device).startswith("cuda")
    if world == 1:
        for j, gi in enumerate(plan[0]):
            full[gi] = local_sketches[j]
    elif on_gpu:
        per_rank = exchange_sketches_device(backend, local_sketches, dist, device, timings)
        for r, idxs in enumerate(plan):
            for j, gi in enumerate(idxs):
                full[gi] = per_rank[r][j]
    else:
        gathered = exchange_sketches([backend.export(s) for s in local_sketches], dist, device)
        for r, idxs in enumerate(plan):
            for j, gi in enumerate(idxs):
                full[gi] = local_sketches[j] if r == rank else backend.import_(gathered[r][j], **(import_params or {}))
    t0 = time.perf_counter()
    kw = dict(query_opts)
    if timings is not None and hasattr(backend, "ctx"):
        kw["stats"] = timings
    hits = np.asarray(backend.query(full, local_sketches, **kw), np.float64).reshape(-1, 5)
    if len(hits):
        hits[:, 0] = np.asarray(mine, np.float64)[hits[:, 0].astype(np.int64)]
    t1 = time.perf_counter()
    if world == 1:
        out = hits[np.lexsort((hits[:, 1], hits[:, 0]))]
    else:
        allh = gather_hits(hits, dist, device)
        out = allh if rank == 0 else None
    if timings is not None:
        timings["query_ms"] = 1e3 * (t1 - t0)
        timings["gather_ms"] = 1e3 * (time.perf_counter() - t1)
        timings["local_hits"] = int(len(hits))
    return out


def all_vs_all(genomes, backend, dist=None, device="cpu", sketch_params=None, query_opts=None, timings=None):
    """genomes: list (identical on every rank) of lists of contigs.  Returns on rank 0 the (n, 5) hit table
    [query, ref, ani, af_query, af_ref] over all ordered pairs; other ranks get None."""
    sketch_params = sketch_params or {}
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    sizes = [sum(len(c) for c in g) for g in genomes]
    plan = partition_by_size(sizes, world)
    mine = plan[rank]
    local = backend.sketch([genomes[i] for i in mine], **sketch_params)
    return query_and_gather(backend, local, mine, plan, dist, device, query_opts, timings, import_params=sketch_params)
