"""Multi-GPU all-vs-all (BASELINE.json configs 3 and 5).

One process per GPU (torch.distributed: NCCL on GPUs, gloo in the CPU tests).  The pair matrix partitions by
query, so the only exchange step is the sketch database:

  1. genomes are split across ranks, balanced by total bases        (partition_by_size)
  2. every rank sketches its share on its own GPU                    (backend.sketch)
  3. sketches are exchanged ONCE.  On GPUs: every rank packs the device arrays of its sketches into one device
     buffer (skb_sketch_pack), NCCL all-gathers those buffers over NVLink, every rank rebuilds the others' sketches
     with one device-to-device copy each (skb_sketch_unpack) - nothing goes through the host and nothing is sorted
     again                                                           (exchange_sketches_device)
     Fallback / CPU tests: an all-gather of the exported host SoA    (exchange_sketches)
  4. every rank builds the full database and queries ITS genomes against it
  5. hits are gathered on rank 0                                     (gather_hits)

There is no collective inside screen / chain / ANI.  `backend` is the object that talks to the device:
CudaBackend (libskb through pyskani_b200.capi) in production; the tests plug a CPU stand-in to exercise the
partitioning / exchange / gather logic under gloo.
"""
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
    """The one data-path collective: every rank contributes its exported sketches, every rank receives all of them
    (list over ranks of lists of export dicts)."""
    header, payload = pack_sketches(local_exports)
    headers = _all_gather_var(header, dist, device)
    payloads = _all_gather_var(payload, dist, device)
    return [unpack_sketches(h, p) for h, p in zip(headers, payloads)]


def exchange_sketches_device(backend, local_sketches, dist, device):
    """Device-resident version of exchange_sketches for backends that can pack/unpack (CudaBackend): returns a list
    over ranks of lists of sketch handles living on this rank's GPU (this rank's entry is `local_sketches` itself)."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    payload, meta = backend.pack(local_sketches, device)              # torch.uint8 cuda tensor, numpy uint8
    metas = _all_gather_var(meta, dist, device)
    n = torch.tensor([payload.numel()], dtype=torch.int64, device=device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(x.item()) for x in sizes]
    m = max(sizes + [16])
    if payload.numel() < m:
        payload = torch.cat([payload, torch.zeros(m - payload.numel(), dtype=torch.uint8, device=device)])
    bufs = torch.empty(world * m, dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(bufs, payload)
    torch.cuda.synchronize(device)
    out = []
    for r in range(world):
        out.append(local_sketches if r == rank else backend.unpack(metas[r], bufs[r * m:r * m + sizes[r]]))
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

    def __init__(self, device_index):
        from . import capi
        self.capi = capi
        self.ctx = capi.Context(device_index)

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

    def query(self, db_sketches, query_sketches, **opts):
        db = self.capi.Database(self.ctx)
        db.add_many(list(db_sketches))
        hits, _ = db.query(query_sketches, **opts)
        return [(h[0], h[1], h[2], h[3], h[4]) for h in hits]


def all_vs_all(genomes, backend, dist=None, device="cpu", sketch_params=None, query_opts=None):
    """genomes: list (identical on every rank) of lists of contigs.  Returns on rank 0 the (n, 5) hit table
    [query, ref, ani, af_query, af_ref] over all ordered pairs; other ranks get None."""
    sketch_params = sketch_params or {}
    query_opts = query_opts or {}
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    sizes = [sum(len(c) for c in g) for g in genomes]
    plan = partition_by_size(sizes, world)
    mine = plan[rank]
    local = backend.sketch([genomes[i] for i in mine], **sketch_params)
    # the full database in global genome order; this rank's own sketches are reused, the others arrive over the fabric
    full = [None] * len(genomes)
    on_gpu = world > 1 and hasattr(backend, "pack") and str(device).startswith("cuda")
    if on_gpu:
        per_rank = exchange_sketches_device(backend, local, dist, device)
        for r, idxs in enumerate(plan):
            for j, gi in enumerate(idxs):
                full[gi] = per_rank[r][j]
    else:
        gathered = exchange_sketches([backend.export(s) for s in local], dist, device) if world > 1 else None
        for r, idxs in enumerate(plan):
            for j, gi in enumerate(idxs):
                full[gi] = local[j] if r == rank else backend.import_(gathered[r][j], **sketch_params)
    hits = backend.query(full, local, **query_opts)
    rows = [(mine[q], r, ani, afq, afr) for (q, r, ani, afq, afr) in hits]
    if world == 1:
        a = np.asarray(rows, np.float64).reshape(-1, 5)
        return a[np.lexsort((a[:, 1], a[:, 0]))]
    allh = gather_hits(rows, dist, device)
    return allh if rank == 0 else None
