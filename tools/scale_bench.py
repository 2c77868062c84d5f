"""Full-size runs of BASELINE.json configs[2..4] on ONE GPU, with the genomes generated on the device.

    python tools/scale_bench.py --families 100 --members 100                       # configs[4]: 10 000 x 10 000 all-vs-all
    python tools/scale_bench.py --families 100 --members 10                        # configs[2]: 1 000 x 1 000 all-vs-all
    python tools/scale_bench.py --families 500 --members 10 --mag-queries 500      # configs[3]: 500 fragmented MAGs vs 5 000
    torchrun --nproc-per-node N tools/scale_bench.py --families 100 --members 100  # all-vs-all over N GPUs: families are dealt
        # to the ranks, every rank sketches its share, the sketches are exchanged once on the device (skb_sketch_pack +
        # NCCL all-gather + skb_sketch_unpack), every rank queries its genomes against the full database

Genomes never exist on the host: a family is a random base genome plus mutated copies (substitutions with probability
0.9 d, indel events with probability 0.1 d, geometric lengths, as pyskani_b200.synth does on the CPU), written as ASCII into
one device buffer that skb_sketch_batch_device reads.  torch is only the random-number generator and the allocator here.
The run checks properties that do not need the oracle: every genome hits itself with ANI 1, every hit stays inside its
family, ANI against the family's base genome decreases with the divergence it was generated with.
"""
import argparse, os, sys, time, json
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyskani_b200 import capi


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--families", type=int, default=20)
    ap.add_argument("--members", type=int, default=10)
    ap.add_argument("--genome-len", type=int, default=5_000_000)
    ap.add_argument("--max-div", type=float, default=0.15)
    ap.add_argument("--query-chunk", type=int, default=500, help="queries per skb_db_query call")
    ap.add_argument("--mag-queries", type=int, default=0, help="configs[3]: N fragmented queries instead of all-vs-all")
    ap.add_argument("--screen-only", action="store_true")
    ap.add_argument("--repeat", type=int, default=1, help="run the query phase N times (the first pays for growing the scratch arena)")
    ap.add_argument("--json", default=None)
    return ap.parse_args()


ACGT = None


def mutate_dev(codes, d, gen):
    """codes: uint8 tensor of 2-bit codes on the device -> mutated copy"""
    n = codes.numel()
    out = codes.clone()
    sub = torch.rand(n, device=codes.device, generator=gen) < 0.9 * d
    out[sub] = (out[sub] + torch.randint(1, 4, (int(sub.sum()),), device=codes.device, generator=gen, dtype=torch.uint8)) & 3
    ev = torch.rand(n, device=codes.device, generator=gen) < 0.1 * d
    idx = ev.nonzero().flatten()
    if idx.numel():
        m = idx.numel()
        u = torch.rand(m, device=codes.device, generator=gen).clamp_(min=1e-9)
        length = torch.clamp((torch.log(u) / np.log(2.0 / 3.0)).long() + 1, min=1, max=50)
        is_ins = torch.rand(m, device=codes.device, generator=gen) < 0.5
        # deletions: drop [p, p + l)
        diff = torch.zeros(n + 64, dtype=torch.int32, device=codes.device)
        dp, dl = idx[~is_ins], length[~is_ins]
        diff.index_add_(0, dp, torch.ones_like(dp, dtype=torch.int32))
        diff.index_add_(0, dp + dl, -torch.ones_like(dp, dtype=torch.int32))
        keep = torch.cumsum(diff[:n], 0) <= 0
        # insertions: l random bases in front of position p
        counts = torch.ones(n, dtype=torch.long, device=codes.device)
        counts[idx[is_ins]] += length[is_ins]
        counts = counts[keep]
        kept = out[keep]
        rep = torch.repeat_interleave(kept, counts)
        orig = torch.zeros(rep.numel(), dtype=torch.bool, device=codes.device)
        orig[torch.cumsum(counts, 0) - 1] = True
        n_new = int((~orig).sum())
        rep[~orig] = torch.randint(0, 4, (n_new,), device=codes.device, generator=gen, dtype=torch.uint8)
        out = rep
    return out


class DeviceBatch:
    """ASCII contigs of several genomes in one device buffer, each contig at a 16-byte aligned offset"""

    def __init__(self, device, cap_bytes):
        self.buf = torch.zeros(cap_bytes, dtype=torch.uint8, device=device)
        self.reset()

    def reset(self):
        self.cur = 64
        self.offs, self.lens, self.gstart = [], [], [0]

    def add_genome(self, contigs):
        for c in contigs:
            l = c.numel()
            assert self.cur + l + 128 <= self.buf.numel(), "device batch buffer too small"
            self.buf[self.cur:self.cur + l] = ACGT[c.long()]
            self.offs.append(self.cur); self.lens.append(l)
            self.cur += (l + 15) // 16 * 16 + 16
        self.gstart.append(len(self.offs))

    def sketch(self, ctx):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sk = ctx.sketch_batch_device(self.buf.data_ptr(), np.array(self.gstart, np.uint32), np.array(self.offs, np.uint64),
                                     np.array(self.lens, np.uint64))
        t1 = time.perf_counter()
        st = ctx.stats()
        bases = int(sum(self.lens))
        self.reset()
        return sk, t1 - t0, st.total_ms / 1e3, bases


def fragment_dev(codes, rng, lo=1000, hi=50000):
    """MAG-style query: log-uniform contig lengths in [lo, hi], shuffled, half reverse-complemented"""
    n = codes.numel()
    cuts, p = [], 0
    while p < n:
        l = int(np.exp(rng.uniform(np.log(lo), np.log(hi))))
        cuts.append((p, min(n, p + l)))
        p += l
    out = []
    for i in rng.permutation(len(cuts)):
        a, b = cuts[i]
        s = codes[a:b]
        if rng.random() < 0.5:
            s = 3 - torch.flip(s, [0])
        out.append(s)
    return out


def main():
    global ACGT
    args = parse()
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        if args.mag_queries:
            raise SystemExit("the MAG configuration is single-GPU in this tool")
    say = print if rank == 0 else (lambda *a, **k: None)

    def tmax(x):          # a per-rank time -> max over ranks
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def tsum(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    ACGT = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    gen = torch.Generator(device=dev); gen.manual_seed(1234 + rank)
    ctx = capi.Context(local)
    F, M, L = args.families, args.members, args.genome_len
    divs = [0.0] + [0.01 + (args.max_div - 0.01) * i / max(1, M - 2) for i in range(M - 1)]
    batch_genomes = max(1, min(100, (600 << 20) // int(L * 1.03)))
    batch = DeviceBatch(dev, int(batch_genomes * (L * 1.03 + 64)) + 4096)

    # context warm-up, timed separately: one full batch through sketch + query grows the context's scratch arena and its
    # first storage slabs (the first few cudaMalloc calls of a process cost 20-100 ms each; later ones ~1 ms)
    t0 = time.perf_counter()
    wbase = torch.randint(0, 4, (L,), device=dev, generator=gen, dtype=torch.uint8)
    for i in range(batch_genomes):
        batch.add_genome([wbase if i % 2 == 0 else torch.roll(wbase, 1000 * i)])
    warm = batch.sketch(ctx)[0]
    wdb = capi.Database(ctx); wdb.add_many(warm); wdb.query(warm[:8])
    del wdb, warm, wbase
    say("context warm-up %.3f s" % (time.perf_counter() - t0), flush=True)

    db = capi.Database(ctx)
    sketches, div_of = [], []
    pending = 0
    t_sketch_wall = t_sketch_dev = 0.0
    total_bases = 0
    bases_kept = {}                      # family -> base codes, only for the families MAG queries are drawn from
    rng = np.random.default_rng(7)
    mag_families = set(rng.choice(F, size=min(F, args.mag_queries), replace=False).tolist()) if args.mag_queries else set()
    t_gen0 = time.perf_counter()
    batch_walls = []

    def flush():
        nonlocal pending, t_sketch_wall, t_sketch_dev, total_bases
        if not pending:
            return
        sk, w, d, b = batch.sketch(ctx)
        t_sketch_wall += w; t_sketch_dev += d; total_bases += b
        batch_walls.append(w)
        sketches.extend(sk)
        pending = 0

    my_fams = list(range(rank, F, world))          # families are dealt round-robin; genome id = family * M + member
    for f in my_fams:
        base = torch.randint(0, 4, (L,), device=dev, generator=gen, dtype=torch.uint8)
        if f in mag_families:
            bases_kept[f] = base
        for m in range(M):
            g = base if m == 0 else mutate_dev(base, divs[m], gen)
            batch.add_genome([g]); div_of.append(divs[m])
            pending += 1
            if pending == batch_genomes:
                flush()
    flush()
    t_gen = time.perf_counter() - t_gen0 - t_sketch_wall
    gid = [f * M + m for f in my_fams for m in range(M)]          # global ids of this rank's sketches
    t_exchange = 0.0
    if world > 1:
        # the one exchange step: packed device buffers, NCCL all-gather over NVLink, device-to-device unpack
        from pyskani_b200 import parallel
        be = parallel.CudaBackend.__new__(parallel.CudaBackend); be.capi = capi; be.ctx = ctx
        # twice: the first exchange pays for the fresh device memory of the block (new slabs from cudaMalloc), the second
        # shows the steady state of a repeated job; then a bare NCCL all-gather of the same bytes for comparison
        ex_log = []
        for rep in range(2):
            per_rank = None
            dist.barrier(); torch.cuda.synchronize()
            tm = {}
            t0 = time.perf_counter()
            per_rank = parallel.exchange_sketches_device(be, sketches, dist, dev, timings=tm)
            t_host = time.perf_counter() - t0
            torch.cuda.synchronize()
            t_exchange = tmax(time.perf_counter() - t0)
            ev = tm.pop("_exchange_events")
            ag_ms = ev[0].elapsed_time(ev[1])
            ex_log.append(dict(total_s=t_exchange, host_s=tmax(t_host), sizes_ms=tmax(tm["exchange_sizes_ms"]), pack_enqueue_ms=tmax(tm["exchange_pack_enqueue_ms"]),
                               adopt_ms=tmax(tm["exchange_adopt_ms"]), allgather_ms=tmax(ag_ms), bytes_in=tm["exchange_bytes_in"], bytes_total=tm["exchange_bytes_total"],
                               collective=tm["exchange_collective"]))
        nbytes = int(ex_log[-1]["bytes_total"]) // world // 256 * 256
        src = torch.empty(nbytes, dtype=torch.uint8, device=dev); dst = torch.empty(nbytes * world, dtype=torch.uint8, device=dev)
        bare = []
        for rep in range(3):
            dist.barrier(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); dist.all_gather_into_tensor(dst, src); e1.record(); torch.cuda.synchronize()
            bare.append(tmax(e0.elapsed_time(e1)))
        del src, dst
        for i, e in enumerate(ex_log):
            say("exchange %d: %.1f ms in all (host until usable %.1f ms: sizes %.2f + pack/enqueue %.2f + adopt %.2f); all-gathers %.2f ms for %.2f GB into "
                "every GPU = %.0f GB/s busbw (%s)" % (i, 1e3 * e["total_s"], 1e3 * e["host_s"], e["sizes_ms"], e["pack_enqueue_ms"], e["adopt_ms"], e["allgather_ms"],
                                                      e["bytes_in"] / 1e9, e["bytes_in"] / e["allgather_ms"] / 1e6, e["collective"]), flush=True)
        say("exchange reference: a bare ncclAllGather of the same %.2f GB per rank: %s ms -> %.0f GB/s busbw" % (
            nbytes / 1e9, ", ".join("%.2f" % b for b in bare), nbytes * (world - 1) / min(bare) / 1e6), flush=True)
        full = [None] * (F * M)
        for r in range(world):
            ids = [f * M + m for f in range(r, F, world) for m in range(M)]
            for j, g in enumerate(ids):
                full[g] = per_rank[r][j]
    else:
        full = sketches
    n = len(full)
    db.add_many(full)
    if batch_walls:
        bw = sorted(batch_walls)
        say("sketch calls: %d, median %.2f ms, slowest %s ms (fresh device memory for the growing database)" % (
            len(bw), 1e3 * bw[len(bw) // 2], ", ".join("%.1f" % (1e3 * x) for x in bw[-4:][::-1])), flush=True)
    total_bases = int(tsum(total_bases)); t_sketch_wall = tmax(t_sketch_wall); t_sketch_dev = tmax(t_sketch_dev)
    say("database: %d genomes on %d GPU(s), %.2f Gbp | generation %.1f s | sketching wall %.3f s (device %.3f s) = %.1f Gbp/s%s" % (
        n, world, total_bases / 1e9, t_gen, t_sketch_wall, t_sketch_dev, total_bases / t_sketch_wall / 1e9,
        " | sketch exchange (pack + NCCL all-gather + unpack) %.3f s" % t_exchange if world > 1 else ""), flush=True)

    # ---- queries
    if args.mag_queries:
        queries, q_family, q_div = [], [], []
        fams = sorted(bases_kept)
        t_q = 0.0
        qb = 0
        for i in range(args.mag_queries):
            f = fams[i % len(fams)]
            d = 0.01 + 0.09 * rng.random()
            g = mutate_dev(bases_kept[f], d, gen)
            batch.add_genome(fragment_dev(g, rng)); q_family.append(f); q_div.append(d)
            pending += 1
            if pending == batch_genomes // 2 or i == args.mag_queries - 1:
                sk, w, dd, b = batch.sketch(ctx)
                queries.extend(sk); t_q += w; qb += b; pending = 0
        print("queries: %d fragmented genomes (%.2f Gbp, %d contigs on average) sketched in %.3f s" % (
            len(queries), qb / 1e9, int(np.mean([s.info().n_contigs for s in queries])), t_q), flush=True)
    else:
        queries, q_family, q_div = sketches, [g // M for g in gid], div_of

    if args.screen_only:
        t0 = time.perf_counter()
        n_pass = 0
        for q0 in range(0, len(queries), args.query_chunk):
            p, _ = db.screen(queries[q0:q0 + args.query_chunk])
            n_pass += int(np.count_nonzero(p))
        t1 = time.perf_counter()
        print("screen only: %d pairs in %.3f s = %.1f M pairs/s, %d pass" % (len(queries) * n, t1 - t0, len(queries) * n / (t1 - t0) / 1e6, n_pass))
        return

    for rep in range(args.repeat):
        if world > 1:
            dist.barrier()
        t_wall = t_screen = t_chain = 0.0
        n_in = 0
        hits = []
        for q0 in range(0, len(queries), args.query_chunk):
            t0 = time.perf_counter()
            h, k = db.query(queries[q0:q0 + args.query_chunk])
            t1 = time.perf_counter()
            st = ctx.stats()
            t_wall += t1 - t0; t_screen += st.screen_ms / 1e3; t_chain += st.chain_ms / 1e3; n_in += k
            hits.extend((q0 + x[0],) + x[1:5] for x in h)
        n_pairs = int(tsum(len(queries))) * n
        t_wall, t_screen, t_chain, n_in_all, n_hits_all = tmax(t_wall), tmax(t_screen), tmax(t_chain), int(tsum(n_in)), int(tsum(len(hits)))
        say("query%s: %d x %d = %.3g pairs in %.3f s wall = %.2f M pairs/s | screen %.3f s (%.1f M pairs/s) | %d pairs chained in %.3f s "
              "(%.0f pairs/s) | %d hits" % (" (repeat %d)" % rep if rep else "", n_pairs // n, n, n_pairs, t_wall, n_pairs / t_wall / 1e6,
                                            t_screen, n_pairs / max(t_screen, 1e-9) / 1e6, n_in_all, t_chain, n_in_all / max(t_chain, 1e-9),
                                            n_hits_all), flush=True)

    # ---- properties
    intra = all(q_family[h[0]] == h[1] // M for h in hits)
    ok_self = True
    if not args.mag_queries:
        self_ani = {h[0]: h[2] for h in hits if gid[h[0]] == h[1]}
        ok_self = len(self_ani) == len(queries) and min(self_ani.values()) > 0.9999
    # ANI against the family's base genome (member 0) must fall as the generating divergence grows
    to_base = {}
    for h in hits:
        if h[1] % M == 0:
            to_base.setdefault(q_family[h[0]], []).append((q_div[h[0]], h[2]))
    mono = True; worst = 0.0
    for f, lst in to_base.items():
        lst.sort()
        for (d0, a0), (d1, a1) in zip(lst, lst[1:]):
            if d1 - d0 > 0.004 and a1 > a0 + 1e-3:
                mono = False
        for d, a in lst:
            if d <= 0.10:
                worst = max(worst, abs((1 - a) - d))
    all_ok = tsum(0.0 if (intra and ok_self and mono) else 1.0) == 0.0
    worst = tmax(worst)
    say("properties%s: hits intra-family %s | self hits ANI 1 %s | ANI monotone in divergence %s | max |(1-ANI) - d| for d<=10%% = %.4f" % (
        " (all ranks)" if world > 1 else "", intra and all_ok, ok_self and all_ok, mono and all_ok, worst))
    if args.json and rank == 0:
        json.dump({"gpus": world, "genomes": n, "queries": n_pairs // n, "bases": total_bases, "sketch_wall_s": t_sketch_wall,
                   "exchange_s": t_exchange, "exchange_log": ex_log if world > 1 else None, "query_wall_s": t_wall, "screen_s": t_screen, "chain_s": t_chain, "pairs": n_pairs,
                   "chained": n_in_all, "hits": n_hits_all, "properties_ok": all_ok, "max_abs_err_vs_divergence": worst},
                  open(args.json, "w"))
    if world > 1:
        dist.barrier(); dist.destroy_process_group()
    if not all_ok:
        raise SystemExit(1)


if __name__ == "__main__":
    main()
