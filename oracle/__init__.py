"""CPU ORACLE — test infrastructure only.

ctypes front-end of oracle/liboracle.so (see oracle/skani_oracle.h for what it restates and why).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs import
this package; the product (pyskani_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class ChainParams(C.Structure):
    _fields_ = [
        ("fragment_length", C.c_int32), ("anchor_score", C.c_double), ("min_anchors", C.c_int32),
        ("min_score", C.c_double), ("max_gap", C.c_double), ("index_band", C.c_int32),
        ("bp_band", C.c_int32), ("frac_cover_cutoff", C.c_double), ("robust", C.c_int32),
        ("median", C.c_int32), ("chunk_mode", C.c_int32), ("order_by_ref", C.c_int32),
        ("count_mode", C.c_int32), ("mean_mode", C.c_int32), ("af_mode", C.c_int32),
        ("af_ext", C.c_double), ("overlap_tol", C.c_double), ("overlap_side", C.c_int32),
        ("switch_mode", C.c_int32), ("denom_mode", C.c_int32), ("min_window_anchors", C.c_int32),
        ("strict_dr", C.c_int32),
    ]


class Result(C.Structure):
    _fields_ = [
        ("ani", C.c_float), ("af_query", C.c_float), ("af_ref", C.c_float),
        ("ani_f64", C.c_double), ("af_query_f64", C.c_double), ("af_ref_f64", C.c_double),
        ("n_anchors", C.c_int64), ("n_windows", C.c_int64), ("n_chains", C.c_int64),
        ("switched", C.c_int32), ("features", C.c_float * 10),
    ]


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = [os.path.join(_HERE, f) for f in ("skani_oracle.cpp", "skani_oracle.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        vp, u64, u32, i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int32
        L.orc_sketch_new.restype = vp
        L.orc_sketch_new.argtypes = [C.POINTER(C.c_char_p), C.POINTER(u64), u32, i32, i32, i32, i32]
        L.orc_sketch_free.argtypes = [vp]
        L.orc_sketch_batch.argtypes = [vp, vp, vp, u32, i32, i32, i32, i32, i32, vp]
        for name, rt in (("orc_sketch_n_seeds", u64), ("orc_sketch_n_markers", u64),
                         ("orc_sketch_n_contigs", u32), ("orc_sketch_total_len", u64)):
            getattr(L, name).restype = rt
            getattr(L, name).argtypes = [vp]
        L.orc_sketch_seeds.argtypes = [vp, vp, vp, vp, vp]
        L.orc_sketch_markers.argtypes = [vp, vp]
        L.orc_sketch_contig_lengths.argtypes = [vp, vp]
        L.orc_screen.restype = i32
        L.orc_screen.argtypes = [vp, vp, C.c_double, i32, C.POINTER(u64)]
        L.orc_chain.argtypes = [vp, vp, C.POINTER(ChainParams), C.POINTER(Result)]
        L.orc_chain_params_default.argtypes = [C.POINTER(ChainParams)]
        L.orc_mm_hash64.restype = u64
        L.orc_mm_hash64.argtypes = [u64]
        L.orc_query.restype = C.c_int64
        L.orc_query.argtypes = [vp, C.POINTER(vp), u64, C.c_double, i32, C.POINTER(ChainParams), i32,
                                vp, vp, C.POINTER(u64)]
        L.orc_query_many.restype = C.c_int64
        L.orc_query_many.argtypes = [C.POINTER(vp), u64, C.POINTER(vp), u64, C.c_double, i32, C.POINTER(ChainParams), i32,
                                     vp, vp, vp, u64, C.POINTER(u64)]
        _LIB = L
    return _LIB


def default_params(**kw):
    p = ChainParams()
    lib().orc_chain_params_default(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


class Sketch:
    """Database::_sketch (reference lib.rs:140-185) on the CPU."""

    def __init__(self, contigs, k=15, c=125, marker_c=1000, seed=True):
        contigs = [bytes(x) if not isinstance(x, bytes) else x for x in contigs]
        n = len(contigs)
        arr = (C.c_char_p * max(n, 1))(*contigs)
        lens = (C.c_uint64 * max(n, 1))(*[len(x) for x in contigs])
        self._h = lib().orc_sketch_new(arr, lens, n, k, c, marker_c, int(seed))
        self.k, self.c, self.marker_c = k, c, marker_c

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_sketch_free(self._h)
            self._h = None

    @property
    def n_seeds(self):
        return lib().orc_sketch_n_seeds(self._h)

    @property
    def n_markers(self):
        return lib().orc_sketch_n_markers(self._h)

    @property
    def total_len(self):
        return lib().orc_sketch_total_len(self._h)

    def seeds(self):
        n = self.n_seeds
        kmer = np.empty(n, np.uint64); pos = np.empty(n, np.uint32)
        contig = np.empty(n, np.uint32); canon = np.empty(n, np.uint8)
        lib().orc_sketch_seeds(self._h, kmer.ctypes.data, pos.ctypes.data, contig.ctypes.data, canon.ctypes.data)
        return kmer, pos, contig, canon

    def markers(self):
        out = np.empty(self.n_markers, np.uint64)
        lib().orc_sketch_markers(self._h, out.ctypes.data)
        return out

    def contig_lengths(self):
        out = np.empty(lib().orc_sketch_n_contigs(self._h), np.uint32)
        lib().orc_sketch_contig_lengths(self._h, out.ctypes.data)
        return out


def sketch_batch(genomes, k=15, c=125, marker_c=1000, seed=True, threads=0):
    """genomes: list of lists of contiguous uint8 numpy arrays / bytes. OpenMP over genomes."""
    flat, starts = [], [0]
    for g in genomes:
        flat.extend(g)
        starts.append(len(flat))
    keep = [np.frombuffer(x, dtype=np.uint8) for x in flat]
    n = len(keep)
    ptrs = (C.c_void_p * max(n, 1))(*[a.ctypes.data for a in keep])
    lens = (C.c_uint64 * max(n, 1))(*[a.size for a in keep])
    gs = (C.c_uint32 * len(starts))(*starts)
    out = (C.c_void_p * max(len(genomes), 1))()
    lib().orc_sketch_batch(ptrs, lens, gs, len(genomes), k, c, marker_c, int(seed), threads, out)
    res = []
    for i in range(len(genomes)):
        s = Sketch.__new__(Sketch)
        s._h, s.k, s.c, s.marker_c = out[i], k, c, marker_c
        res.append(s)
    return res


def screen(query, ref, screen_val=0.8, rescue_small=True):
    shared = C.c_uint64(0)
    ok = lib().orc_screen(query._h, ref._h, screen_val, int(rescue_small), C.byref(shared))
    return bool(ok), shared.value


def chain(ref, query, params=None, **kw):
    p = params if params is not None else default_params(**kw)
    r = Result()
    lib().orc_chain(ref._h, query._h, C.byref(p), C.byref(r))
    return r


def query(q, refs, screen_val=0.8, rescue_small=True, params=None, threads=1):
    """pyskani Database.query loop (reference lib.rs:616-657). Returns (hit indices, results, n screened-in)."""
    p = params if params is not None else default_params()
    n = len(refs)
    hs = (C.c_void_p * max(n, 1))(*[r._h for r in refs])
    idx = np.empty(max(n, 1), np.uint32)
    res = (Result * max(n, 1))()
    ns = C.c_uint64(0)
    nh = lib().orc_query(q._h, hs, n, screen_val, int(rescue_small), C.byref(p), threads,
                         idx.ctypes.data, C.addressof(res), C.byref(ns))
    return idx[:nh].copy(), [res[i] for i in range(nh)], ns.value


def query_many(queries, refs, screen_val=0.8, rescue_small=True, params=None, threads=0):
    """pyskani's query loop for many queries, parallel over (query, ref) pairs.
    Returns (hit_q, hit_r, results, n screened-in), hits ordered by (query, ref)."""
    p = params if params is not None else default_params()
    nq, nr = len(queries), len(refs)
    qh = (C.c_void_p * max(nq, 1))(*[q._h for q in queries])
    rh = (C.c_void_p * max(nr, 1))(*[r._h for r in refs])
    cap = max(1024, 16 * nq)
    while True:
        hq = np.empty(cap, np.uint32); hr = np.empty(cap, np.uint32)
        res = (Result * cap)()
        ns = C.c_uint64(0)
        nh = lib().orc_query_many(qh, nq, rh, nr, screen_val, int(rescue_small), C.byref(p), threads,
                                  hq.ctypes.data, hr.ctypes.data, C.addressof(res), cap, C.byref(ns))
        if nh <= cap:
            return hq[:nh].copy(), hr[:nh].copy(), [res[i] for i in range(nh)], ns.value
        cap = nh


class Gbdt:
    """CPU evaluator of a gbdt-rs regression ensemble (crate gbdt 0.1.3, Cargo.lock:541-542), restated from its
    published algorithm: prediction = bias + sum over the first `iterations` trees of shrinkage * tree(x), every product and
    sum rounded to f32 (gbdt-rs' ValueType) in tree order; a tree descends left when x[feature] < feature_value, follows
    `missing` when the feature is f32::MIN, and returns `pred` at a leaf.  Reads the serde_json text itself (python json),
    independently of the product's C++ parser."""
    UNKNOWN = np.float32(-3.40282347e+38)

    def __init__(self, text):
        import json
        d = json.loads(text)
        conf = d["conf"]
        self.n_features = int(conf["feature_size"])
        self.shrinkage = np.float32(conf["shrinkage"])
        self.bias = np.float32(d.get("bias", 0.0))
        self.trees = []
        for t in d["trees"][:int(conf.get("iterations", len(d["trees"])))]:
            self.trees.append([(int(n["value"]["feature_index"]), np.float32(n["value"]["feature_value"]), np.float32(n["value"]["pred"]),
                                int(n["value"].get("missing", 0)), bool(n["value"]["is_leaf"]), int(n["left"]), int(n["right"]))
                               for n in t["tree"]["tree"]])

    def tree(self, nodes, x):
        i = 0
        while True:
            f, thr, pred, miss, leaf, l, r = nodes[i]
            if leaf:
                return pred
            v = np.float32(x[f])
            if v == self.UNKNOWN:
                if miss == 0:
                    return pred
                nxt = l if miss < 0 else r
            else:
                nxt = l if v < thr else r
            if nxt == 0:
                return pred
            i = nxt

    def predict(self, x):
        p = self.bias
        for nodes in self.trees:
            p = np.float32(p + np.float32(self.shrinkage * self.tree(nodes, x)))
        return p


LEARNED_MIN_COV = 150000.0


def learned_ani(result, model, robust=False, median=False):
    """ANI of an orc_chain result after the learned correction, as f32 (what Hit.identity carries)."""
    if robust or median or not (result.ani > 0) or result.features[9] < LEARNED_MIN_COV:
        return np.float32(result.ani)
    a = float(model.predict([result.features[i] for i in range(10)])) / 100.0
    return np.float32(min(1.0, max(0.0, a)))


def mm_hash64(x):
    return lib().orc_mm_hash64(x)
