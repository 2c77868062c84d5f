"""ctypes binding of include/skb.h (libskb.so).

This is the thinnest possible Python view of the C ABI: it is what bench.py and the parity tests call,
and what the reference's maintainers would mirror as an `extern "C"` block on the Rust side
(INTEGRATION.md).  It never computes anything itself and has no CPU fallback: if libskb.so is missing,
or no CUDA device is usable, it raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SKB_LIB") or os.path.join(_HERE, "libskb.so")     # SKB_LIB: kernel-variant builds (tools/seed_variants.sh)

SKB_OK, SKB_ERR_ARG, SKB_ERR_CUDA, SKB_ERR_NOMEM, SKB_ERR_KEY, SKB_ERR_UNSUPPORTED = range(6)


class SketchParams(C.Structure):
    _fields_ = [("k", C.c_int32), ("c", C.c_int32), ("marker_c", C.c_int32)]


class QueryOpts(C.Structure):
    _fields_ = [("cutoff", C.c_double), ("learned_ani", C.c_int32), ("median", C.c_int32),
                ("robust", C.c_int32), ("faster_small", C.c_int32)]


class Hit(C.Structure):
    _fields_ = [("query_index", C.c_uint32), ("ref_index", C.c_uint32), ("ani", C.c_float),
                ("af_query", C.c_float), ("af_ref", C.c_float), ("n_windows", C.c_uint32),
                ("n_chains", C.c_uint32), ("n_anchors", C.c_uint32)]


HIT_DTYPE = np.dtype([('query_index', '<u4'), ('ref_index', '<u4'), ('ani', '<f4'), ('af_query', '<f4'), ('af_ref', '<f4'),
                      ('n_windows', '<u4'), ('n_chains', '<u4'), ('n_anchors', '<u4')])


class SketchInfo(C.Structure):
    _fields_ = [("n_seeds", C.c_uint64), ("n_markers", C.c_uint64), ("total_len", C.c_uint64),
                ("n_contigs", C.c_uint32), ("k", C.c_int32), ("c", C.c_int32), ("marker_c", C.c_int32),
                ("has_seeds", C.c_int32), ("reference_only", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("h2d_ms", C.c_float), ("seed_ms", C.c_float), ("index_ms", C.c_float),
                ("screen_ms", C.c_float), ("chain_ms", C.c_float), ("total_ms", C.c_float),
                ("kernels_launched", C.c_uint64), ("h2d_raw_bytes", C.c_uint64), ("h2d_packed_bytes", C.c_uint64)]


# every symbol include/skb.h declares (tests/test_abi.py checks the library exports all of them)
SYMBOLS = [
    "skb_ctx_create", "skb_ctx_destroy", "skb_last_error", "skb_ctx_stats", "skb_ctx_sync", "skb_ctx_stream",
    "skb_ctx_set_host_threads", "skb_ctx_set_priority",
    "skb_host_alloc", "skb_host_free", "skb_dev_alloc", "skb_dev_free", "skb_memcpy_h2d",
    "skb_sketch_batch", "skb_sketch_batch_device", "skb_sketch_free", "skb_sketch_free_many", "skb_sketch_info", "skb_sketch_export",
    "skb_sketch_import", "skb_sketch_pack_size", "skb_sketch_pack", "skb_sketch_unpack",
    "skb_exchange_segment_size", "skb_exchange_create", "skb_exchange_ptr", "skb_exchange_pack", "skb_exchange_order_after",
    "skb_exchange_adopt",
    "skb_exchange_free",
    "skb_model_load_json", "skb_model_free", "skb_model_info", "skb_model_predict", "skb_db_set_model", "skb_db_create", "skb_db_destroy", "skb_db_add", "skb_db_add_many", "skb_db_replace", "skb_db_size", "skb_db_query",
    "skb_hits_free", "skb_db_screen", "skb_version",
]

_lib = None


def handle_array(sketches):
    """(count, pointer usable as `skb_sketch_t* const*`, object to keep alive) for a SketchArray or any sequence of Sketch"""
    if isinstance(sketches, SketchArray):
        return len(sketches), sketches.handles.ctypes.data, sketches
    arr = np.fromiter((s._h for s in sketches), np.uint64, len(sketches))
    return len(arr), arr.ctypes.data, arr


class SkbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libskb error {code}: {msg}")
        self.code = code


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(pyskani_b200 has no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        vp, u64, u32, i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int32
        L.skb_version.restype = C.c_char_p
        L.skb_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
        L.skb_ctx_destroy.argtypes = [vp]
        L.skb_last_error.restype = C.c_char_p
        L.skb_last_error.argtypes = [vp]
        L.skb_ctx_stats.argtypes = [vp, C.POINTER(Stats)]
        L.skb_ctx_sync.argtypes = [vp]
        L.skb_ctx_stream.restype = vp
        L.skb_ctx_stream.argtypes = [vp]
        L.skb_ctx_set_host_threads.argtypes = [vp, i32]
        L.skb_ctx_set_priority.argtypes = [vp, i32]
        L.skb_host_alloc.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
        L.skb_host_free.argtypes = [vp, vp]
        L.skb_dev_alloc.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
        L.skb_dev_free.argtypes = [vp, vp]
        L.skb_memcpy_h2d.argtypes = [vp, vp, vp, C.c_size_t]
        L.skb_sketch_batch.argtypes = [vp, C.POINTER(SketchParams), i32, u32, vp, vp, vp, vp]
        L.skb_sketch_batch_device.argtypes = [vp, C.POINTER(SketchParams), i32, u32, vp, vp, vp, vp, vp]
        L.skb_sketch_free.argtypes = [vp]
        L.skb_sketch_free_many.argtypes = [u32, vp]
        L.skb_sketch_info.argtypes = [vp, C.POINTER(SketchInfo)]
        L.skb_sketch_export.argtypes = [vp, vp, vp, vp, vp, vp, vp]
        L.skb_sketch_import.argtypes = [vp, C.POINTER(SketchParams), i32, u64, vp, vp, vp, vp, u64, vp, u32, vp,
                                        C.POINTER(vp)]
        L.skb_sketch_pack_size.argtypes = [u32, vp, C.POINTER(u64), C.POINTER(u64)]
        L.skb_sketch_pack.argtypes = [vp, u32, vp, vp, u64, vp, u64]
        L.skb_sketch_unpack.argtypes = [vp, vp, u64, vp, u64, vp, u32, C.POINTER(u32)]
        L.skb_exchange_segment_size.argtypes = [u32, vp, i32, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64)]
        L.skb_exchange_order_after.argtypes = [vp, vp, i32]
        L.skb_exchange_create.argtypes = [vp, u64, C.POINTER(vp)]
        L.skb_exchange_ptr.restype = vp
        L.skb_exchange_ptr.argtypes = [vp]
        L.skb_exchange_pack.argtypes = [vp, u64, u64, u32, vp, i32]
        L.skb_exchange_adopt.argtypes = [vp, u32, vp, vp, vp, vp, u32, vp]
        L.skb_exchange_free.argtypes = [vp]
        L.skb_model_load_json.argtypes = [vp, C.c_char_p, C.c_size_t, C.POINTER(vp)]
        L.skb_model_free.argtypes = [vp]
        L.skb_model_info.argtypes = [vp, C.POINTER(u32), C.POINTER(u32), C.POINTER(u32)]
        L.skb_model_predict.argtypes = [vp, vp, u32, u32, vp]
        L.skb_db_set_model.argtypes = [vp, vp]
        L.skb_db_create.argtypes = [vp, C.POINTER(vp)]
        L.skb_db_destroy.argtypes = [vp]
        L.skb_db_add.argtypes = [vp, vp, C.POINTER(u32)]
        L.skb_db_add_many.argtypes = [vp, u32, vp, C.POINTER(u32)]
        L.skb_db_replace.argtypes = [vp, u32, vp]
        L.skb_db_size.restype = u64
        L.skb_db_size.argtypes = [vp]
        L.skb_db_query.argtypes = [vp, u32, vp, C.POINTER(QueryOpts), C.POINTER(C.POINTER(Hit)), C.POINTER(u64),
                                   C.POINTER(u64)]
        L.skb_hits_free.argtypes = [C.POINTER(Hit)]
        L.skb_db_screen.argtypes = [vp, u32, vp, C.c_double, i32, vp, vp]
        _lib = L
    return _lib


class Context:
    def __init__(self, device=0):
        self._h = C.c_void_p()
        rc = lib().skb_ctx_create(device, C.byref(self._h))
        if rc != SKB_OK:
            self._h = None
            raise SkbError(rc, f"cannot create a CUDA context on device {device} (pyskani_b200 has no CPU fallback)")

    def check(self, rc):
        if rc != SKB_OK:
            raise SkbError(rc, lib().skb_last_error(self._h).decode())

    def close(self):
        if self._h:
            try:
                lib().skb_ctx_destroy(self._h)
            except (TypeError, ImportError, AttributeError):    # interpreter shutdown: the library may already be gone      # interpreter shutdown: module globals are already gone
                pass
            self._h = None

    def __del__(self):
        # sketches and databases hold their own reference to the device state; destroying the handle is safe
        self.close()

    def stats(self):
        s = Stats()
        self.check(lib().skb_ctx_stats(self._h, C.byref(s)))
        return s

    def sync(self):
        self.check(lib().skb_ctx_sync(self._h))

    @property
    def stream(self):
        return lib().skb_ctx_stream(self._h)

    def set_priority(self, high=True):
        """Highest stream priority for this context's kernels: those of other contexts of the device only fill its gaps."""
        self.check(lib().skb_ctx_set_priority(self._h, 1 if high else 0))

    def set_host_threads(self, n):
        """Threads of the host ingest pipeline of sketch calls on host buffers (0/1: all bytes travel as ASCII; < 0: default)."""
        self.check(lib().skb_ctx_set_host_threads(self._h, int(n)))

    # ---- sketching
    def sketch_batch(self, genomes, k=15, c=125, marker_c=1000, seed=True):
        """genomes: list of lists of bytes-like contigs (host memory). Returns a list of Sketch."""
        flat, starts = [], [0]
        for g in genomes:
            flat.extend(g)
            starts.append(len(flat))
        n = len(flat)
        keep = [np.frombuffer(x, dtype=np.uint8) if len(x) else np.zeros(0, np.uint8) for x in flat]
        ptrs = (C.c_void_p * max(n, 1))(*[a.ctypes.data if a.size else None for a in keep])
        lens = (C.c_uint64 * max(n, 1))(*[a.size for a in keep])
        gs = (C.c_uint32 * len(starts))(*starts)
        out = np.zeros(max(len(genomes), 1), np.uint64)
        p = SketchParams(k, c, marker_c)
        self.check(lib().skb_sketch_batch(self._h, C.byref(p), int(seed), len(genomes), gs, ptrs, lens, out.ctypes.data))
        return SketchArray(self, out[:len(genomes)])

    def sketch_batch_device(self, seq_dev_ptr, genome_contig_start, contig_offsets, contig_lens,
                            k=15, c=125, marker_c=1000, seed=True):
        gs = np.ascontiguousarray(genome_contig_start, np.uint32)
        offs = np.ascontiguousarray(contig_offsets, np.uint64)
        lens = np.ascontiguousarray(contig_lens, np.uint64)
        ng = len(gs) - 1
        out = np.zeros(max(ng, 1), np.uint64)
        p = SketchParams(k, c, marker_c)
        self.check(lib().skb_sketch_batch_device(self._h, C.byref(p), int(seed), ng, gs.ctypes.data, seq_dev_ptr,
                                                 offs.ctypes.data, lens.ctypes.data, out.ctypes.data))
        return SketchArray(self, out[:ng])

    def import_sketch(self, kmer, pos, contig, canonical, markers, contig_lengths, k=15, c=125, marker_c=1000,
                      has_seeds=True):
        kmer = np.ascontiguousarray(kmer, np.uint64); pos = np.ascontiguousarray(pos, np.uint32)
        contig = np.ascontiguousarray(contig, np.uint32); canonical = np.ascontiguousarray(canonical, np.uint8)
        markers = np.ascontiguousarray(markers, np.uint64); cl = np.ascontiguousarray(contig_lengths, np.uint32)
        out = C.c_void_p()
        p = SketchParams(k, c, marker_c)
        self.check(lib().skb_sketch_import(self._h, C.byref(p), int(has_seeds), len(kmer), kmer.ctypes.data,
                                           pos.ctypes.data, contig.ctypes.data, canonical.ctypes.data, len(markers),
                                           markers.ctypes.data, len(cl), cl.ctypes.data, C.byref(out)))
        return Sketch(self, out.value)

    # ---- device-to-device transfer (multi-GPU exchange)
    def pack_size(self, sketches):
        """(payload bytes on the device, descriptor bytes on the host) that pack() needs for these sketches"""
        n, hs, _keep = handle_array(sketches)
        pb, mb = C.c_uint64(), C.c_uint64()
        self.check(lib().skb_sketch_pack_size(n, hs, C.byref(pb), C.byref(mb)))
        return pb.value, mb.value

    def pack(self, sketches, payload_dev_ptr, payload_bytes):
        """Concatenates the sketches' device arrays into the device buffer; returns the host descriptor (uint8 array)."""
        n, hs, _keep = handle_array(sketches)
        _, mb = self.pack_size(sketches)
        meta = np.zeros(mb, np.uint8)
        self.check(lib().skb_sketch_pack(self._h, n, hs, payload_dev_ptr, payload_bytes, meta.ctypes.data, mb))
        return meta

    def unpack(self, meta, payload_dev_ptr, payload_bytes):
        """Rebuilds the sketches described by `meta` from a device payload (one device-to-device copy)."""
        meta = np.ascontiguousarray(meta, np.uint8)
        n = int(meta[4:8].view(np.uint32)[0]) if meta.size >= 8 else 0
        out = np.zeros(max(n, 1), np.uint64)
        got = C.c_uint32()
        self.check(lib().skb_sketch_unpack(self._h, meta.ctypes.data, meta.size, payload_dev_ptr, payload_bytes, out.ctypes.data, n,
                                           C.byref(got)))
        return SketchArray(self, out[:got.value])

    def segment_size(self, sketches, reference_only=False):
        """(head segment bytes, body segment bytes, descriptor bytes) of these sketches inside an exchange block"""
        n, hs, _keep = handle_array(sketches)
        hb, bb, mb = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self.check(lib().skb_exchange_segment_size(n, hs, int(reference_only), C.byref(hb), C.byref(bb), C.byref(mb)))
        return hb.value, bb.value, mb.value

    def exchange(self, nbytes):
        return Exchange(self, nbytes)

    def host_alloc(self, nbytes):
        p = C.c_void_p()
        self.check(lib().skb_host_alloc(self._h, nbytes, C.byref(p)))
        return p.value

    def host_free(self, p):
        lib().skb_host_free(self._h, p)

    def dev_alloc(self, nbytes):
        p = C.c_void_p()
        self.check(lib().skb_dev_alloc(self._h, nbytes, C.byref(p)))
        return p.value

    def dev_free(self, p):
        lib().skb_dev_free(self._h, p)

    def memcpy_h2d(self, dst, src, nbytes):
        self.check(lib().skb_memcpy_h2d(self._h, dst, src, nbytes))


class Model:
    """A gbdt-rs regression ensemble on the device (skani's learned-ANI model, include/skb.h)."""

    def __init__(self, ctx, json_text):
        self.ctx = ctx
        self._h = C.c_void_p()
        raw = json_text.encode() if isinstance(json_text, str) else bytes(json_text)
        ctx.check(lib().skb_model_load_json(ctx._h, raw, len(raw), C.byref(self._h)))

    def info(self):
        a, b, c = C.c_uint32(), C.c_uint32(), C.c_uint32()
        self.ctx.check(lib().skb_model_info(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return {"n_trees": a.value, "n_nodes": b.value, "n_features": c.value}

    def predict(self, rows):
        rows = np.ascontiguousarray(rows, np.float32)
        if rows.ndim == 1:
            rows = rows[None, :]
        out = np.empty(rows.shape[0], np.float32)
        self.ctx.check(lib().skb_model_predict(self._h, rows.ctypes.data, rows.shape[0], rows.shape[1], out.ctypes.data))
        return out

    def __del__(self):
        if getattr(self, "_h", None):
            try:
                lib().skb_model_free(self._h)
            except (TypeError, ImportError, AttributeError):    # interpreter shutdown: the library may already be gone
                pass
            self._h = None


class Exchange:
    """One block of sketch storage holding a segment per rank (include/skb.h, "zero-copy exchange region")."""

    def __init__(self, ctx, nbytes):
        self.ctx, self.nbytes = ctx, int(nbytes)
        self._h = C.c_void_p()
        ctx.check(lib().skb_exchange_create(ctx._h, self.nbytes, C.byref(self._h)))

    @property
    def ptr(self):
        return lib().skb_exchange_ptr(self._h)

    def pack(self, head_offset, body_offset, sketches, reference_only=False):
        n, hs, _keep = handle_array(sketches)
        self.ctx.check(lib().skb_exchange_pack(self._h, int(head_offset), int(body_offset), n, hs, int(reference_only)))

    def order_after(self, cuda_stream, bodies):
        """orders the context behind work enqueued on another CUDA stream (raw handle): heads now, bodies lazily"""
        self.ctx.check(lib().skb_exchange_order_after(self._h, C.c_void_p(int(cuda_stream)), int(bodies)))

    def adopt(self, head_offsets, body_offsets, meta_bytes, max_sketches):
        """-> list (one per peer) of lists of Sketch whose arrays are views into this block"""
        ns = len(head_offsets)
        ho = (C.c_uint64 * max(ns, 1))(*[int(o) for o in head_offsets])
        bo = (C.c_uint64 * max(ns, 1))(*[int(o) for o in body_offsets])
        mbs = (C.c_uint64 * max(ns, 1))(*[int(m) for m in meta_bytes])
        out = np.zeros(max(max_sketches, 1), np.uint64)
        counts = (C.c_uint32 * max(ns, 1))()
        self.ctx.check(lib().skb_exchange_adopt(self._h, ns, ho, bo, mbs, out.ctypes.data, max_sketches, counts))
        res, k = [], 0
        for i in range(ns):
            res.append(SketchArray(self.ctx, out[k:k + counts[i]]))
            k += counts[i]
        return res

    def close(self):
        if getattr(self, "_h", None):
            try:
                lib().skb_exchange_free(self._h)
            except (TypeError, ImportError, AttributeError):    # interpreter shutdown: the library may already be gone
                pass
            self._h = None

    __del__ = close


class Sketch:
    def __init__(self, ctx, handle, owner=None):
        """owner: the SketchArray the handle belongs to (it frees the handle); None = this object frees it"""
        self.ctx, self._h, self._owner = ctx, int(handle) if handle else None, owner

    def __del__(self):
        if getattr(self, "_h", None) and self._owner is None:
            try:
                lib().skb_sketch_free(self._h)
            except (TypeError, ImportError, AttributeError):    # interpreter shutdown: the library may already be gone
                pass
        self._h = None

    def info(self):
        i = SketchInfo()
        self.ctx.check(lib().skb_sketch_info(self._h, C.byref(i)))
        return i

    def export(self):
        i = self.info()
        kmer = np.empty(i.n_seeds, np.uint64); pos = np.empty(i.n_seeds, np.uint32)
        contig = np.empty(i.n_seeds, np.uint32); canon = np.empty(i.n_seeds, np.uint8)
        markers = np.empty(i.n_markers, np.uint64); cl = np.empty(i.n_contigs, np.uint32)
        self.ctx.check(lib().skb_sketch_export(self._h, kmer.ctypes.data, pos.ctypes.data, contig.ctypes.data,
                                               canon.ctypes.data, markers.ctypes.data, cl.ctypes.data))
        return dict(kmer=kmer, pos=pos, contig=contig, canonical=canon, markers=markers, contig_lengths=cl)


class SketchArray:
    """The sketches one call produced: a numpy array of handles that is passed to the C ABI as it is and freed with ONE call
    (skb_sketch_free_many) - a per-handle Python object for each of 1 000 sketches per step costs more than the query.
    Behaves like a list of Sketch (indexing, slicing, iteration, len, +); items and slices are views that keep it alive."""

    def __init__(self, ctx, handles, parents=None):
        self.ctx = ctx
        self.handles = np.ascontiguousarray(handles, np.uint64)
        self._parents = parents            # None: this array owns its handles; else the owning arrays it is a view of

    def __len__(self):
        return len(self.handles)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return SketchArray(self.ctx, self.handles[i], parents=(self,))
        return Sketch(self.ctx, self.handles[i], owner=self)

    def __iter__(self):
        return (Sketch(self.ctx, h, owner=self) for h in self.handles)

    def __add__(self, other):
        other = other if isinstance(other, SketchArray) else SketchArray.of(self.ctx, other)
        return SketchArray(self.ctx, np.concatenate([self.handles, other.handles]), parents=(self, other))

    __iadd__ = __add__

    @staticmethod
    def concat(ctx, arrays):
        arrays = [a if isinstance(a, SketchArray) else SketchArray.of(ctx, a) for a in arrays]
        h = np.concatenate([a.handles for a in arrays]) if arrays else np.zeros(0, np.uint64)
        return SketchArray(ctx, h, parents=tuple(arrays))

    @staticmethod
    def gather(ctx, n, parts):
        """n handles in a given order: parts = [(index list, SketchArray), ...]; every index appears exactly once"""
        h = np.zeros(n, np.uint64)
        for idxs, arr in parts:
            h[np.asarray(idxs, np.int64)] = arr.handles
        return SketchArray(ctx, h, parents=tuple(a for _, a in parts))

    @staticmethod
    def of(ctx, sketches):
        """a (non-owning) array over any sequence of Sketch / SketchArray items"""
        if isinstance(sketches, SketchArray):
            return sketches
        items = list(sketches)
        return SketchArray(ctx, np.fromiter((s._h for s in items), np.uint64, len(items)), parents=tuple(items))

    def __del__(self):
        if getattr(self, "_parents", 0) is None and len(self.handles):
            try:
                lib().skb_sketch_free_many(len(self.handles), self.handles.ctypes.data)
            except (TypeError, ImportError, AttributeError):    # interpreter shutdown: the library may already be gone
                pass
        self.handles = np.zeros(0, np.uint64)


class Database:
    def __init__(self, ctx):
        self.ctx = ctx
        self._h = C.c_void_p()
        ctx.check(lib().skb_db_create(ctx._h, C.byref(self._h)))
        self._keep = []

    def __del__(self):
        if getattr(self, "_h", None):
            try:
                lib().skb_db_destroy(self._h)
            except (TypeError, ImportError, AttributeError):    # interpreter shutdown: the library may already be gone
                pass
            self._h = None

    def add(self, sketch):
        idx = C.c_uint32()
        self.ctx.check(lib().skb_db_add(self._h, sketch._h, C.byref(idx)))
        self._keep.append(sketch)
        return idx.value

    def add_many(self, sketches):
        n, hs, keep = handle_array(sketches)
        idx = C.c_uint32()
        self.ctx.check(lib().skb_db_add_many(self._h, n, hs, C.byref(idx)))
        self._keep.append(sketches)
        return idx.value

    def set_model(self, model):
        self.ctx.check(lib().skb_db_set_model(self._h, model._h if model is not None else None))
        self._model = model

    def __len__(self):
        return lib().skb_db_size(self._h)

    def query_array(self, queries, cutoff=0.0, learned_ani=0, median=False, robust=False, faster_small=False):
        """skb_db_query; hits as a numpy structured array (HIT_DTYPE) plus the number of screened-in pairs"""
        n, qs, _keep = handle_array(queries)
        o = QueryOpts(cutoff, learned_ani, int(median), int(robust), int(faster_small))
        hits = C.POINTER(Hit)()
        nh, ns = C.c_uint64(0), C.c_uint64(0)
        self.ctx.check(lib().skb_db_query(self._h, n, qs, C.byref(o), C.byref(hits), C.byref(nh), C.byref(ns)))
        out = np.zeros(0, HIT_DTYPE)
        if nh.value:
            raw = np.ctypeslib.as_array(C.cast(hits, C.POINTER(C.c_uint8)), shape=(nh.value * C.sizeof(Hit),))
            out = raw.view(HIT_DTYPE).copy()
        lib().skb_hits_free(hits)
        return out, ns.value

    def query(self, queries, cutoff=0.0, learned_ani=0, median=False, robust=False, faster_small=False):
        n, qs, _keep = handle_array(queries)
        o = QueryOpts(cutoff, learned_ani, int(median), int(robust), int(faster_small))
        hits = C.POINTER(Hit)()
        nh, ns = C.c_uint64(0), C.c_uint64(0)
        self.ctx.check(lib().skb_db_query(self._h, n, qs, C.byref(o), C.byref(hits), C.byref(nh), C.byref(ns)))
        out = []
        if nh.value:
            raw = np.ctypeslib.as_array(C.cast(hits, C.POINTER(C.c_uint8)), shape=(nh.value * C.sizeof(Hit),))
            out = raw.view(HIT_DTYPE).tolist()       # tuples in skb_hit_t field order
        lib().skb_hits_free(hits)
        return out, ns.value

    def screen(self, queries, cutoff=0.8, rescue_small=True):
        nr = len(self)
        n, qs, _keep = handle_array(queries)
        ok = np.zeros((n, nr), np.uint8); shared = np.zeros((n, nr), np.uint32)
        self.ctx.check(lib().skb_db_screen(self._h, n, qs, cutoff, int(rescue_small), ok.ctypes.data, shared.ctypes.data))
        return ok.astype(bool), shared
