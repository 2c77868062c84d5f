"""Multi-GPU all-vs-all (BASELINE.json configs 3 and 5).

One process per GPU (torch.distributed: NCCL on GPUs, gloo in the CPU tests).  The pair matrix partitions by
query, so the only exchange step is the sketch database:

  1. genomes are split across ranks, balanced by total bases        (partition_by_size)
  2. every rank sketches its share on its own GPU                    (backend.sketch)
  3. sketches are exchanged ONCE, device to device, without staging  (exchange_sketches_device):
       a. a 4-number all-gather tells every rank the segment sizes of every other rank
       b. every rank allocates ONE block of sketch storage with a head segment (descriptor, marker sets) and a body
          segment (seed arrays) per rank and packs its own sketches into its two segments (skb_exchange_pack: one
          gather kernel); peers get what a database member needs, not the query-side arrays (half of the bytes)
       c. two collectives fill the other segments in place - heads first, then the 25x larger bodies: ncclAllGather
          straight into the block when the segments are (nearly) the same size, a group of per-rank broadcasts on exact
          sizes otherwise
       d. the peers' sketches are adopted as views into the block as soon as the heads are there (skb_exchange_adopt:
          no unpack copy); libskb waits for the bodies only when it first reads seed arrays, so the marker screen of
          step 4 overlaps the body transfer
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


def segment_layout(seg_bytes, uniform_slack=0.10, base=0):
    """Offsets of the per-rank segments of one region of the exchange block, starting at `base`.  Segments of (nearly)
    equal size are laid out at a uniform stride so that ONE in-place ncclAllGather can fill the region; otherwise they are
    packed back to back and a group of exact-size broadcasts is used.  Returns (offsets, end of the region, stride or 0)."""
    world = len(seg_bytes)
    stride = max(list(seg_bytes) + [256])
    if stride * world <= (1.0 + uniform_slack) * max(sum(seg_bytes), 1):
        return [base + r * stride for r in range(world)], base + stride * world, stride
    offs, cur = [], base
    for s in seg_bytes:
        offs.append(cur)
        cur += s
    return offs, max(cur, base + 256), 0


def _gather_region(block, offs, sizes, stride, rank, dist):
    """one collective that fills the peers' segments of a region in place"""
    world = len(offs)
    if stride:
        dist.all_gather_into_tensor(block[offs[0]:offs[0] + stride * world], block[offs[rank]:offs[rank] + stride])
    else:
        views = [block[offs[r]:offs[r] + sizes[r]] for r in range(world)]
        dist.all_gather(views, views[rank])                      # exact sizes: grouped broadcasts


def exchange_sketches_device(backend, local_sketches, dist, device, timings=None, reference_only=True):
    """The one data-path exchange, device resident: returns a list over ranks of lists of sketch handles living on this
    rank's GPU (this rank's entry is `local_sketches` itself; the others are views into one exchange block).
      1. a 4-number all-gather tells every rank the segment sizes of every other rank
      2. every rank packs its sketches into its HEAD segment (descriptor, marker sets) and BODY segment (seed arrays)
      3. collective 1 fills the heads (small), collective 2 the bodies (large); the peers' sketches are adopted as soon
         as the heads are there, and libskb waits for the bodies only when seed arrays are first read, so the marker
         screen of the following query runs while the bodies are still on the wire
    reference_only: peers receive what a database member needs (half of the bytes), not what a query needs.
    `timings` (dict, optional) receives the phase milliseconds and the byte counts."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    ctx = backend.ctx
    t0 = time.perf_counter()
    head, body, mb = ctx.segment_size(local_sketches, reference_only)
    mine = torch.tensor([head, body, mb, len(local_sketches)], dtype=torch.int64, device=device)
    allsz = torch.empty(4 * world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(allsz, mine)
    sz = allsz.tolist()
    heads, bodies, metas, counts = sz[0::4], sz[1::4], sz[2::4], sz[3::4]
    h_offs, h_end, h_stride = segment_layout(heads)
    b_offs, total, b_stride = segment_layout(bodies, base=h_end)
    ex = ctx.exchange(total)
    t1 = time.perf_counter()
    ex.pack(h_offs[rank], b_offs[rank], local_sketches, reference_only)        # asynchronous on libskb's stream
    lib_stream = torch.cuda.ExternalStream(ctx.stream, device=device)
    cur = torch.cuda.current_stream(device)
    ev_in, ev_a, ev_b = torch.cuda.Event(), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev_in.record(lib_stream)
    cur.wait_event(ev_in)                                     # the collectives start when this rank's segments are packed
    block = torch.as_tensor(_DevBlock(ex.ptr, total), device=device)
    ev_a.record(cur)
    _gather_region(block, h_offs, heads, h_stride, rank, dist)
    ex.order_after(cur.cuda_stream, bodies=False)             # libskb may read the heads from here on
    _gather_region(block, b_offs, bodies, b_stride, rank, dist)
    ev_b.record(cur)
    ex.order_after(cur.cuda_stream, bodies=True)              # ... and the bodies when it first needs seed arrays
    t2 = time.perf_counter()
    peers = [r for r in range(world) if r != rank]
    adopted = ex.adopt([h_offs[r] for r in peers], [b_offs[r] for r in peers], [metas[r] for r in peers],
                       sum(counts[r] for r in peers))
    t3 = time.perf_counter()
    ex.close()                                                # the block now belongs to the adopted sketches
    out = [None] * world
    out[rank] = local_sketches
    for r, lst in zip(peers, adopted):
        out[r] = lst
    if timings is not None:
        timings["exchange_sizes_ms"] = 1e3 * (t1 - t0)
        timings["exchange_pack_enqueue_ms"] = 1e3 * (t2 - t1)
        timings["exchange_adopt_ms"] = 1e3 * (t3 - t2)
        timings["exchange_ms"] = 1e3 * (t3 - t0)              # host time until the database is usable; bodies may still fly
        timings["_exchange_events"] = (ev_a, ev_b)            # all-gather time is read after the step's final sync
        timings["exchange_bytes_in"] = int(sum(heads) + sum(bodies) - heads[rank] - bodies[rank])
        timings["exchange_bytes_total"] = int(sum(heads) + sum(bodies))
        timings["exchange_collective"] = "%s (heads) + %s (bodies)" % (
            "ncclAllGather in place" if h_stride else "grouped ncclBroadcast", "ncclAllGather in place" if b_stride else "grouped ncclBroadcast")
    return out


def sort_hits(table):
    """rows ordered by (query, ref).  Every rank's rows already are, and no query is shared between ranks, so a stable sort
    on the query column restores the global order."""
    return table[np.argsort(table[:, 0], kind="stable")] if len(table) else table


_GATHER_ROWS = {}       # world size -> rows every rank sends per gather (grows when a rank had more)
_GATHER_BUFS = {}       # (world size, capacity, device) -> pinned (send, receive) blocks


def gather_hits(local_hits, dist, device="cpu", sort=True):
    """hits: (n, 5) float64 rows [query_global, ref_global, ani, af_query, af_ref]; returned on every rank, ordered by
    (query, ref) unless sort=False (then rank after rank; the reference's own hit order is arbitrary, lib.rs:640).
    ONE collective and one host round trip in the steady state: every rank sends a fixed-capacity block whose first row holds
    its row count; the capacity is remembered per world size and, if some rank had more rows than fit, all ranks see that in
    the counts and repeat the collective with a larger capacity (every rank takes the same decision from the same data)."""
    import torch
    world = dist.get_world_size()
    rows = np.ascontiguousarray(np.asarray(local_hits, np.float64).reshape(-1, 5))
    cap = _GATHER_ROWS.get(world, 1024)
    on_gpu = str(device).startswith("cuda")
    while True:
        if on_gpu:
            # pinned host blocks (kept per world size and capacity): the small copies on either side of the collective then
            # run at link speed instead of being staged through pageable memory
            key = (world, cap, str(device))
            bufs = _GATHER_BUFS.get(key)
            if bufs is None:
                _GATHER_BUFS.clear()
                bufs = (torch.zeros((cap + 1, 5), dtype=torch.float64).pin_memory(),
                        torch.zeros((world * (cap + 1), 5), dtype=torch.float64).pin_memory())
                _GATHER_BUFS[key] = bufs
            h_send, h_recv = bufs
            block = h_send.numpy()
        else:
            block = np.zeros((cap + 1, 5), np.float64)
        block[0, 0] = len(rows)
        k = min(len(rows), cap)
        block[1:1 + k] = rows[:k]
        send = h_send.to(device, non_blocking=True) if on_gpu else torch.from_numpy(block)
        allr = torch.empty((world * (cap + 1), 5), dtype=torch.float64, device=device)
        dist.all_gather_into_tensor(allr, send)
        if on_gpu:
            h_recv.copy_(allr, non_blocking=True)
            torch.cuda.current_stream(device).synchronize()
            allr = h_recv.numpy().reshape(world, cap + 1, 5)
        else:
            allr = allr.numpy().reshape(world, cap + 1, 5)
        counts = [int(allr[r, 0, 0]) for r in range(world)]
        if max(counts) <= cap:
            break
        cap = 1 << int(np.ceil(np.log2(max(counts) * 1.25)))
        _GATHER_ROWS[world] = cap
    allh = np.concatenate([allr[r, 1:1 + counts[r]] for r in range(world)]) if world else np.zeros((0, 5))
    return sort_hits(allh) if sort else allh


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
        db.add_many(db_sketches)
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
                     import_params=None, sort=True):
    """Steps 3-5 for sketches that already exist on this rank: exchange, query this rank's genomes against the full
    database, gather the hit table [query, ref, ani, af_query, af_ref] (global ids; ordered by (query, ref) unless
    sort=False) on rank 0."""
    query_opts = query_opts or {}
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    n_total = sum(len(p) for p in plan)
    full = [None] * n_total
    on_gpu = world > 1 and hasattr(backend, "ctx") and str(device).startswith("cuda")
    if hasattr(backend, "ctx"):
        # handle arrays, not per-sketch Python objects: the database of a step is assembled by index arithmetic
        arr = backend.capi.SketchArray
        local_sketches = arr.of(backend.ctx, local_sketches)
        per_rank = exchange_sketches_device(backend, local_sketches, dist, device, timings) if world > 1 else [local_sketches]
        full = arr.gather(backend.ctx, n_total, [(plan[r], per_rank[r]) for r in range(world)])
    elif world == 1:
        for j, gi in enumerate(plan[0]):
            full[gi] = local_sketches[j]
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
        out = sort_hits(hits) if sort else hits
    else:
        allh = gather_hits(hits, dist, device, sort)
        out = allh if rank == 0 else None
    if timings is not None:
        timings["query_ms"] = 1e3 * (t1 - t0)
        timings["gather_ms"] = 1e3 * (time.perf_counter() - t1)
        timings["local_hits"] = int(len(hits))
        ev = timings.pop("_exchange_events", None)
        if ev is not None:
            ev[1].synchronize()
            timings["exchange_allgather_ms"] = ev[0].elapsed_time(ev[1])
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


def all_vs_all_pipelined(sketch_batches, query_backend, query_opts=None, timings=None, sort=True):
    """Single-GPU all-vs-all whose queries run UNDER the host->device ingest of the following sketch batches.

    sketch_batches: an iterable that yields the sketches of one batch after the other (capi.SketchArray or lists of Sketch;
    typically a generator whose body calls skb_sketch_batch on host buffers through the SKETCHING context - the caller's
    thread is inside that call, with the GIL released, most of the time).  query_backend: a CudaBackend on a SECOND context
    of the same device (own stream, own scratch), so that its kernels fill the gaps the GPU has while PCIe and the host's
    compaction threads set the pace.  A pair (i, j) is computable as soon as both genomes are sketched; when batch k arrives
    a worker thread runs
        queries = batches 0..k-1, database = batch k        and        queries = batch k, database = batches 0..k
    which covers every ordered pair exactly once.  Only the two calls of the LAST batch are exposed after the ingest ends.
    Returns the (n, 5) table [query, ref, ani, af_query, af_ref] with global indices in arrival order of the genomes
    (ordered by (query, ref) unless sort=False)."""
    import queue
    import threading
    query_opts = dict(query_opts or {})
    arr = query_backend.capi.SketchArray
    q = queue.Queue()
    rows, err = [], []
    stats = {}

    def work():
        try:
            done, base = [], 0
            while True:
                part = q.get()
                if part is None:
                    return
                part = arr.of(query_backend.ctx, part)
                n = len(part)
                if n == 0:
                    continue
                if done:
                    old = arr.concat(query_backend.ctx, done)
                    h = np.asarray(query_backend.query(part, old, stats=stats, **query_opts), np.float64).reshape(-1, 5)
                    if len(h):
                        h[:, 1] += base
                        rows.append(h)
                done.append(part)
                full = arr.concat(query_backend.ctx, done)
                h = np.asarray(query_backend.query(full, part, stats=stats, **query_opts), np.float64).reshape(-1, 5)
                if len(h):
                    h[:, 0] += base
                    rows.append(h)
                base += n
        except BaseException as e:          # surfaced on the caller's thread
            err.append(e)

    t = threading.Thread(target=work, name="skb-query")
    t.start()
    t0 = time.perf_counter()
    try:
        for part in sketch_batches:
            q.put(part)
            if err:
                break
    finally:
        t1 = time.perf_counter()
        q.put(None)
        t.join()
    t2 = time.perf_counter()
    if err:
        raise err[0]
    table = np.concatenate(rows) if rows else np.zeros((0, 5))
    if timings is not None:
        timings.update(stats)
        timings["sketch_ms"] = 1e3 * (t1 - t0)          # the ingest-bound part; queries of earlier batches ran under it
        timings["query_ms"] = 1e3 * (t2 - t1)           # what was left of the queries once the last batch was sketched
        timings["exchange_ms"] = 0.0
        timings["gather_ms"] = 0.0
        timings["local_hits"] = int(len(table))
    return sort_hits(table) if sort else table
