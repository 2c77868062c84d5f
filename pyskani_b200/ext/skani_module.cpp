// skani_module.cpp — the `pyskani_b200._skani` extension: pyskani's Python API on top of libskb's C ABI.
//
// Mirrors, class by class and argument by argument, the PyO3 module of the reference
// (src/pyskani/_skani/lib.rs:230-758, hit.rs:23-105, sketch.rs:16-32; stubs in src/pyskani/_skani.pyi):
//   Hit(identity, query_name, query_fraction, reference_name, reference_fraction)
//   Sketch.name / .c / .amino_acid
//   Database(path=None, *, compression=125, marker_compression=1000, k=15, format=None)
//   Database.load / .open / .sketch / .query / .save / .flush / .path / .compression / .marker_compression,
//   context-manager protocol.
// All numeric work happens in libskb.so (CUDA); this file holds only argument handling, naming, the three
// storage back-ends of lib.rs:42-123 and their bincode layout.  No CPU fallback.
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <algorithm>
#include <cerrno>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <stdexcept>
#include <string>
#include <sys/stat.h>
#include <unordered_map>
#include <vector>

#include "../../include/skb.h"

namespace py = pybind11;

namespace {

// ------------------------------------------------------------------------------------------- errors
struct OsError { int code; std::string msg; };

[[noreturn]] void throw_os(int code, const std::string& msg) { throw OsError{code, msg}; }

void check(skb_ctx_t* ctx, int rc) {
    if (rc == SKB_OK) return;
    std::string msg = ctx ? skb_last_error(ctx) : "libskb call failed";
    switch (rc) {
        case SKB_ERR_ARG: throw py::value_error(msg);
        case SKB_ERR_KEY: throw py::key_error(msg);
        case SKB_ERR_NOMEM: throw std::bad_alloc();
        default: throw std::runtime_error(msg);
    }
}

// one CUDA context per process, created on first use (device: $PYSKANI_B200_DEVICE, default 0)
skb_ctx_t* global_ctx() {
    static skb_ctx_t* ctx = nullptr;
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    if (!ctx) {
        int dev = 0;
        if (const char* e = std::getenv("PYSKANI_B200_DEVICE")) dev = std::atoi(e);
        int rc = skb_ctx_create(dev, &ctx);
        if (rc != SKB_OK) {
            ctx = nullptr;
            throw std::runtime_error("pyskani_b200 needs a CUDA device (sm_100a); no usable device " + std::to_string(dev) +
                                     " was found and there is no CPU fallback");
        }
    }
    return ctx;
}

// ------------------------------------------------------------------------------------------- bincode 1.3 (default options)
struct Writer {
    std::string buf;
    void u8(uint8_t v) { buf.push_back((char)v); }
    void u32(uint32_t v) { buf.append((const char*)&v, 4); }
    void u64(uint64_t v) { buf.append((const char*)&v, 8); }
    void str(const std::string& s) { u64(s.size()); buf.append(s); }
};

struct Reader {
    const uint8_t* p; size_t n, off = 0;
    void need(size_t k) { if (off + k > n) throw py::value_error("io error: unexpected end of file"); }   // bincode's message for a short read
    uint8_t u8() { need(1); return p[off++]; }
    uint32_t u32() { need(4); uint32_t v; std::memcpy(&v, p + off, 4); off += 4; return v; }
    uint64_t u64() { need(8); uint64_t v; std::memcpy(&v, p + off, 8); off += 8; return v; }
    std::string str() { uint64_t l = u64(); need(l); std::string s((const char*)p + off, l); off += l; return s; }
    bool boolean() { uint8_t v = u8(); if (v > 1) throw py::value_error("invalid value: expected a bool"); return v != 0; }
};

struct Params { uint64_t c = 125, k = 15, marker_c = 1000; };

// Host copy of one sketch: what skani::types::Sketch serialises (A.1), in flat arrays
struct HostSketch {
    std::string file_name;
    bool has_seeds = true;
    std::vector<uint64_t> kmer; std::vector<uint32_t> pos, contig; std::vector<uint8_t> canonical;   // sorted by (kmer, contig, pos)
    std::vector<std::string> contigs;
    uint64_t total_len = 0;
    std::vector<uint32_t> contig_lengths;
    std::vector<uint64_t> markers;
    Params params;
};

const char* CODON_LETTERS = "KNKNTTTTRSRSIIMIQHQHPPPPRRRRLLLLEDEDAAAAGGGGVVVV*Y*YSSSS*CWCLFLF";   // codon (A=0,C=1,G=2,T=3) -> amino acid

// SketchParams { c, k, marker_c, use_syncs, use_aa, acgt_to_aa_encoding: Vec<u64>, acgt_to_aa_letters: Vec<u8>, orf_size }
void write_params(Writer& w, const Params& p) {
    w.u64(p.c); w.u64(p.k); w.u64(p.marker_c);
    w.u8(0); w.u8(0);
    w.u64(64);
    for (int i = 0; i < 64; i++) { const char* a = std::strchr("ACDEFGHIKLMNPQRSTVWY*", CODON_LETTERS[i]); w.u64((uint64_t)(a - "ACDEFGHIKLMNPQRSTVWY*")); }
    w.u64(64);
    for (int i = 0; i < 64; i++) w.u8((uint8_t)CODON_LETTERS[i]);
    w.u64(30);   // ORF_SIZE
}
Params read_params(Reader& r) {
    Params p;
    p.c = r.u64(); p.k = r.u64(); p.marker_c = r.u64();
    r.boolean(); r.boolean();
    uint64_t n = r.u64(); for (uint64_t i = 0; i < n; i++) r.u64();
    n = r.u64(); for (uint64_t i = 0; i < n; i++) r.u8();
    r.u64();
    return p;
}

// Sketch { file_name, kmer_seeds_k: Option<HashMap<u64, SmallVec<[SeedPosition;1]>>>, contigs, total_sequence_length,
//          contig_lengths, repetitive_kmers, marker_seeds: HashSet<u64>, marker_c, c, k, contig_order, amino_acid }
// SeedPosition { pos: u32, canonical: bool, contig_index: u32, phase: u8 }
void write_sketch(Writer& w, const HostSketch& s, bool markers_only) {
    w.str(s.file_name);
    if (markers_only || !s.has_seeds) {
        w.u8(0);
    } else {
        w.u8(1);
        uint64_t distinct = 0;
        for (size_t i = 0; i < s.kmer.size(); i++) if (i == 0 || s.kmer[i] != s.kmer[i - 1]) distinct++;
        w.u64(distinct);
        for (size_t i = 0; i < s.kmer.size();) {
            size_t j = i;
            while (j < s.kmer.size() && s.kmer[j] == s.kmer[i]) j++;
            w.u64(s.kmer[i]); w.u64(j - i);
            for (size_t t = i; t < j; t++) { w.u32(s.pos[t]); w.u8(s.canonical[t]); w.u32(s.contig[t]); w.u8(0); }
            i = j;
        }
    }
    w.u64(s.contigs.size());
    for (auto& c : s.contigs) w.str(c);
    w.u64(s.total_len);
    w.u64(s.contig_lengths.size());
    for (auto l : s.contig_lengths) w.u32(l);
    w.u64(0);   // repetitive_kmers (unused since skani 0.3.0, lib.rs:177-182)
    w.u64(s.markers.size());
    for (auto m : s.markers) w.u64(m);
    w.u64(s.params.marker_c); w.u64(s.params.c); w.u64(s.params.k);
    w.u64(0);   // contig_order
    w.u8(0);    // amino_acid
}
HostSketch read_sketch(Reader& r) {
    HostSketch s;
    s.file_name = r.str();
    uint8_t tag = r.u8();
    if (tag > 1) throw py::value_error("invalid value: expected an Option tag");
    s.has_seeds = tag == 1;
    if (tag) {
        uint64_t n = r.u64();
        for (uint64_t i = 0; i < n; i++) {
            uint64_t km = r.u64(), m = r.u64();
            for (uint64_t t = 0; t < m; t++) {
                uint32_t pos = r.u32(); bool canon = r.boolean(); uint32_t ci = r.u32(); r.u8();
                s.kmer.push_back(km); s.pos.push_back(pos); s.contig.push_back(ci); s.canonical.push_back(canon);
            }
        }
    }
    uint64_t nc = r.u64();
    for (uint64_t i = 0; i < nc; i++) s.contigs.push_back(r.str());
    s.total_len = r.u64();
    uint64_t nl = r.u64();
    for (uint64_t i = 0; i < nl; i++) s.contig_lengths.push_back(r.u32());
    r.u64();
    uint64_t nm = r.u64();
    for (uint64_t i = 0; i < nm; i++) s.markers.push_back(r.u64());
    s.params.marker_c = r.u64(); s.params.c = r.u64(); s.params.k = r.u64();
    r.u64();
    r.boolean();
    return s;
}

// ------------------------------------------------------------------------------------------- files (utils.rs:25-72)
bool exists(const std::string& p) { struct stat st; return ::stat(p.c_str(), &st) == 0; }

void mkdirs(const std::string& path) {
    std::string cur;
    for (size_t i = 0; i <= path.size(); i++) {
        if (i == path.size() || path[i] == '/') {
            if (!cur.empty() && !exists(cur) && ::mkdir(cur.c_str(), 0777) != 0 && errno != EEXIST)
                throw_os(errno, "Failed to create " + path);
        }
        if (i < path.size()) cur.push_back(path[i]);
    }
}
std::string read_file(const std::string& path) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) throw_os(errno, "Failed to open " + path);
    std::string out; char buf[1 << 16]; size_t n;
    while ((n = std::fread(buf, 1, sizeof buf, f)) > 0) out.append(buf, n);
    std::fclose(f);
    return out;
}
void write_file(const std::string& path, const std::string& data, const char* mode) {
    FILE* f = std::fopen(path.c_str(), mode);
    if (!f) throw_os(errno, "Failed to create " + path);
    if (!data.empty() && std::fwrite(data.data(), 1, data.size(), f) != data.size()) { int e = errno; std::fclose(f); throw_os(e, "Failed to write " + path); }
    std::fclose(f);
}
uint64_t file_size(const std::string& path) { struct stat st; return ::stat(path.c_str(), &st) == 0 ? (uint64_t)st.st_size : 0; }
std::string join(const std::string& a, const std::string& b) { return a.empty() || a.back() == '/' ? a + b : a + "/" + b; }

// str / bytes / bytearray / any buffer -> borrowed byte view (utils.rs:74-103)
struct View { py::buffer_info info; py::object keep; const uint8_t* ptr; uint64_t len; };
View view_of(const py::handle& h) {
    View v;
    if (py::isinstance<py::str>(h)) {
        Py_ssize_t n = 0;
        const char* p = PyUnicode_AsUTF8AndSize(h.ptr(), &n);
        if (!p) throw py::error_already_set();
        v.keep = py::reinterpret_borrow<py::object>(h); v.ptr = (const uint8_t*)p; v.len = (uint64_t)n;
        return v;
    }
    if (!PyObject_CheckBuffer(h.ptr())) throw py::type_error("expected str, bytes, bytearray or an object supporting the buffer protocol");
    py::buffer b = py::reinterpret_borrow<py::buffer>(h);
    v.info = b.request();
    if (v.info.ndim > 1) throw py::value_error("expected a one-dimensional buffer");
    if (v.info.ndim == 1 && v.info.strides[0] != v.info.itemsize) throw py::value_error("expected a contiguous buffer");
    v.keep = py::reinterpret_borrow<py::object>(h);
    v.ptr = (const uint8_t*)v.info.ptr; v.len = (uint64_t)(v.info.size * v.info.itemsize);
    return v;
}

std::string fsdecode(const py::handle& h) { return py::module_::import("os").attr("fsdecode")(h).cast<std::string>(); }

std::string basename(const std::string& p) {   // Path::file_name (lib.rs:629-635)
    size_t e = p.find_last_not_of('/');
    if (e == std::string::npos) return p;
    size_t s = p.find_last_of('/', e);
    return p.substr(s == std::string::npos ? 0 : s + 1, e - (s == std::string::npos ? 0 : s + 1) + 1);
}

// ------------------------------------------------------------------------------------------- Hit (hit.rs)
struct Hit {
    float identity; std::string query_name; float query_fraction; std::string reference_name; float reference_fraction;
};

// ------------------------------------------------------------------------------------------- Sketch (sketch.rs)
struct SketchHandle {
    skb_sketch_t* h = nullptr;
    ~SketchHandle() { if (h) skb_sketch_free(h); }
};
struct Sketch {
    std::shared_ptr<SketchHandle> handle;
    std::string name;
    std::vector<std::string> contig_names;
    uint64_t c = 125;
    bool amino_acid = false;
};

HostSketch export_sketch(const Sketch& s, const Params& params, bool markers_only = false) {
    skb_sketch_info_t info;
    check(global_ctx(), skb_sketch_info(s.handle->h, &info));
    HostSketch o;
    o.file_name = s.name; o.has_seeds = info.has_seeds != 0; o.params = params;
    o.markers.resize(info.n_markers); o.contig_lengths.resize(info.n_contigs);
    if (markers_only) {
        // markers.bin holds no seeds (get_markers_only, lib.rs:495): leave the seed arrays on the device
        check(global_ctx(), skb_sketch_export(s.handle->h, nullptr, nullptr, nullptr, nullptr, o.markers.data(), o.contig_lengths.data()));
    } else {
        o.kmer.resize(info.n_seeds); o.pos.resize(info.n_seeds); o.contig.resize(info.n_seeds); o.canonical.resize(info.n_seeds);
        check(global_ctx(), skb_sketch_export(s.handle->h, o.kmer.data(), o.pos.data(), o.contig.data(), o.canonical.data(),
                                              o.markers.data(), o.contig_lengths.data()));
    }
    o.contigs = s.contig_names; o.total_len = info.total_len;
    return o;
}

Sketch import_sketch(const HostSketch& hs) {
    skb_sketch_params_t p{(int32_t)hs.params.k, (int32_t)hs.params.c, (int32_t)hs.params.marker_c};
    Sketch s;
    s.handle = std::make_shared<SketchHandle>();
    check(global_ctx(), skb_sketch_import(global_ctx(), &p, hs.has_seeds ? 1 : 0, hs.kmer.size(), hs.kmer.data(), hs.pos.data(),
                                          hs.contig.data(), hs.canonical.data(), hs.markers.size(), hs.markers.data(),
                                          (uint32_t)hs.contig_lengths.size(), hs.contig_lengths.data(), &s.handle->h));
    s.name = hs.file_name; s.contig_names = hs.contigs; s.c = hs.params.c;
    return s;
}

// ------------------------------------------------------------------------------------------- learned-ANI model
// skani::regression::get_model (lib.rs:614) deserialises a gbdt-rs ensemble embedded in the skani crate.  Those weights are
// not part of pyskani's sources; here the same JSON comes from a file: Database(model=...), Database.set_model(...) or
// $PYSKANI_B200_MODEL.
struct ModelHandle {
    skb_model_t* h = nullptr;
    ~ModelHandle() { if (h) skb_model_free(h); }
};
std::shared_ptr<ModelHandle> load_model_file(const std::string& path) {
    const std::string text = read_file(path);
    auto m = std::make_shared<ModelHandle>();
    check(global_ctx(), skb_model_load_json(global_ctx(), text.data(), text.size(), &m->h));
    return m;
}

// ------------------------------------------------------------------------------------------- Database (lib.rs:132-741)
enum class Storage { Memory, Folder, Consolidated };

struct IndexEntry { std::string file_name; uint64_t offset, length; };

struct Database {
    Params params;
    Storage storage = Storage::Memory;
    std::string folder;
    std::vector<Sketch> items;                          // index == index inside the device database
    std::unordered_map<std::string, size_t> by_name;    // Memory keys / Consolidated index keys
    std::vector<IndexEntry> index;                      // Consolidated: entries in append order (= offset order)
    skb_db_t* db = nullptr;
    std::shared_ptr<ModelHandle> model;
    void set_model(std::shared_ptr<ModelHandle> m) {
        check(global_ctx(), skb_db_set_model(db, m ? m->h : nullptr));
        model = std::move(m);
    }
    void model_from(const py::object& arg) {      // explicit argument, else the environment
        if (!arg.is_none()) { set_model(load_model_file(fsdecode(arg))); return; }
        if (const char* e = std::getenv("PYSKANI_B200_MODEL")) if (*e) set_model(load_model_file(e));
    }
    // Readers (query) share, writers (sketch, flush) exclude: the reference's RwLocks (lib.rs:135-136).  NEVER taken while
    // the GIL is held: a thread that blocks on `mu` with the GIL would stop the holder from ever re-acquiring the GIL.
    std::shared_mutex mu;

    Database() { check(global_ctx(), skb_db_create(global_ctx(), &db)); }
    ~Database() { if (db) skb_db_destroy(db); }
    Database(const Database&) = delete;

    // DatabaseStorage::store (lib.rs:49-91)
    void store(const Sketch& s) {
        if (storage == Storage::Memory) return;
        Writer w;
        write_params(w, params);
        write_sketch(w, export_sketch(s, params), false);
        if (storage == Storage::Folder) {
            write_file(join(folder, s.name + ".sketch"), w.buf, "wb");
        } else {
            if (by_name.count(s.name)) throw py::value_error("duplicate name in sketches: \"" + s.name + "\"");   // lib.rs:67-72
            const std::string path = join(folder, "sketches.db");
            IndexEntry e{s.name, file_size(path), (uint64_t)w.buf.size()};
            write_file(path, w.buf, "ab");
            index.push_back(e);
        }
    }

    void add(Sketch s, bool persist) {
        if (persist) store(s);
        auto it = by_name.find(s.name);
        if (it != by_name.end()) {
            // same name again (Memory / Folder storage): the reference's sketch store is a map keyed by name, so the new
            // sketch takes the place of the old one and a query reports the name once (lib.rs:45,64,617-640)
            check(global_ctx(), skb_db_replace(db, (uint32_t)it->second, s.handle->h));
            items[it->second] = std::move(s);
            return;
        }
        uint32_t idx = 0;
        check(global_ctx(), skb_db_add(db, s.handle->h, &idx));
        by_name[s.name] = items.size();
        items.push_back(std::move(s));
    }

    // Database::_save_markers (lib.rs:187-201): (SketchParams, Vec<Sketch markers-only>)
    void save_markers(const std::string& path) {
        Writer w;
        write_params(w, params);
        w.u64(items.size());
        for (auto& s : items) write_sketch(w, export_sketch(s, params, true), true);
        write_file(path, w.buf, "wb");
    }
    // Database::_save_index (lib.rs:203-215)
    static void save_index(const std::string& path, std::vector<IndexEntry> idx) {
        std::sort(idx.begin(), idx.end(), [](const IndexEntry& a, const IndexEntry& b) { return a.offset < b.offset; });
        Writer w;
        w.u64(idx.size());
        for (auto& e : idx) { w.str(e.file_name); w.u64(e.offset); w.u64(e.length); }
        write_file(path, w.buf, "wb");
    }
    // Database::_flush (lib.rs:217-227)
    void flush() {          // callers hold no GIL
        std::unique_lock<std::shared_mutex> lk(mu);
        if (storage == Storage::Memory) return;
        save_markers(join(folder, "markers.bin"));
        if (storage == Storage::Consolidated) save_index(join(folder, "index.db"), index);
    }
};

Storage parse_format(const py::object& format) {
    if (format.is_none()) return Storage::Consolidated;
    std::string f = format.cast<std::string>();
    if (f == "consolidated") return Storage::Consolidated;
    if (f == "separated") return Storage::Folder;
    throw py::value_error("invalid format: " + f);
}

// Database.open / Database.load (lib.rs:251-337).  Both put every sketch into HBM: the reference's lazy
// per-query disk reads (lib.rs:99-119) make no sense next to 180 GB of device memory; `open` keeps the storage
// mode of the folder so that later sketch() calls append to it, `load` detaches into a Memory database.
std::unique_ptr<Database> open_impl(const py::object& path, bool detach, const py::object& model) {
    const std::string folder = fsdecode(path);
    auto db = std::make_unique<Database>();
    db->model_from(model);
    std::string raw = read_file(join(folder, "markers.bin"));
    Reader r{(const uint8_t*)raw.data(), raw.size()};
    db->params = read_params(r);
    uint64_t n = r.u64();
    std::vector<std::string> names;
    for (uint64_t i = 0; i < n; i++) names.push_back(read_sketch(r).file_name);
    const bool consolidated = exists(join(folder, "index.db")) && exists(join(folder, "sketches.db"));
    std::unordered_map<std::string, IndexEntry> idx;
    std::string blob;
    if (consolidated) {
        std::string iraw = read_file(join(folder, "index.db"));
        Reader ir{(const uint8_t*)iraw.data(), iraw.size()};
        uint64_t ne = ir.u64();
        for (uint64_t i = 0; i < ne; i++) { IndexEntry e; e.file_name = ir.str(); e.offset = ir.u64(); e.length = ir.u64(); idx[e.file_name] = e; db->index.push_back(e); }
        blob = read_file(join(folder, "sketches.db"));
    }
    for (auto& name : names) {
        HostSketch hs;
        if (consolidated) {
            auto it = idx.find(name);
            if (it == idx.end()) throw py::key_error(name);
            if (it->second.offset + it->second.length > blob.size()) throw py::value_error("io error: unexpected end of file");
            Reader sr{(const uint8_t*)blob.data() + it->second.offset, it->second.length};
            read_params(sr);
            hs = read_sketch(sr);
        } else {
            std::string sraw = read_file(join(folder, name + ".sketch"));
            Reader sr{(const uint8_t*)sraw.data(), sraw.size()};
            read_params(sr);
            hs = read_sketch(sr);
        }
        hs.params = db->params;
        db->add(import_sketch(hs), false);
    }
    if (detach) { db->storage = Storage::Memory; db->index.clear(); }
    else { db->storage = consolidated ? Storage::Consolidated : Storage::Folder; db->folder = folder; }
    return db;
}

// Database::_sketch (lib.rs:140-185) through skb_sketch_batch
Sketch sketch_impl(Database& db, const std::string& name, const py::tuple& contigs, bool seed) {
    std::vector<View> views;
    views.reserve(contigs.size());
    for (auto item : contigs) views.push_back(view_of(item));
    std::vector<const uint8_t*> ptrs; std::vector<uint64_t> lens;
    Sketch s;
    s.name = name; s.c = db.params.c;
    for (size_t i = 0; i < views.size(); i++) {
        ptrs.push_back(views[i].ptr); lens.push_back(views[i].len);
        if (views[i].len >= SKB_MIN_LENGTH_CONTIG) s.contig_names.push_back(name + "_" + std::to_string(i));   // lib.rs:157
    }
    uint32_t gstart[2] = {0, (uint32_t)views.size()};
    skb_sketch_params_t p{(int32_t)db.params.k, (int32_t)db.params.c, (int32_t)db.params.marker_c};
    s.handle = std::make_shared<SketchHandle>();
    int rc;
    {
        py::gil_scoped_release nogil;   // lib.rs:493 / 569
        rc = skb_sketch_batch(global_ctx(), &p, seed ? 1 : 0, 1, gstart, ptrs.data(), lens.data(), &s.handle->h);
    }
    check(global_ctx(), rc);
    return s;
}

// Batched form of sketch_impl: items = [(name, contigs), ...]; ONE skb_sketch_batch call for all genomes.
std::vector<Sketch> sketch_many_impl(Database& db, const py::sequence& items, bool seed) {
    std::vector<View> views;
    std::vector<const uint8_t*> ptrs; std::vector<uint64_t> lens; std::vector<uint32_t> gstart{0};
    std::vector<Sketch> out;
    for (auto item : items) {
        py::tuple tp = py::cast<py::tuple>(item);
        if (tp.size() != 2) throw py::value_error("expected (name, contigs) pairs");
        Sketch s;
        s.name = tp[0].cast<std::string>(); s.c = db.params.c;
        py::object contigs = tp[1];
        if (py::isinstance<py::str>(contigs) || py::isinstance<py::bytes>(contigs) || PyByteArray_Check(contigs.ptr()) || PyMemoryView_Check(contigs.ptr()))
            contigs = py::make_tuple(contigs);      // a single contig
        size_t i = 0;
        for (auto cobj : contigs) {
            views.push_back(view_of(cobj));
            ptrs.push_back(views.back().ptr); lens.push_back(views.back().len);
            if (views.back().len >= SKB_MIN_LENGTH_CONTIG) s.contig_names.push_back(s.name + "_" + std::to_string(i));
            i++;
        }
        gstart.push_back((uint32_t)ptrs.size());
        out.push_back(std::move(s));
    }
    const uint32_t n = (uint32_t)out.size();
    std::vector<skb_sketch_t*> handles(n, nullptr);
    skb_sketch_params_t p{(int32_t)db.params.k, (int32_t)db.params.c, (int32_t)db.params.marker_c};
    int rc = SKB_OK;
    {
        // one GPU batch holds fewer than 2^31 bases: a long list goes down in slices of about 1.25 G bases (whole genomes)
        py::gil_scoped_release nogil;
        constexpr uint64_t SLICE_BASES = 1250ull << 20;
        for (uint32_t g0 = 0; g0 < n && rc == SKB_OK;) {
            uint32_t g1 = g0;
            uint64_t bases = 0;
            while (g1 < n) {
                uint64_t b = 0;
                for (uint32_t ci = gstart[g1]; ci < gstart[g1 + 1]; ci++) b += lens[ci];
                if (g1 > g0 && bases + b > SLICE_BASES) break;
                bases += b; g1++;
            }
            std::vector<uint32_t> gs(g1 - g0 + 1);
            for (uint32_t g = g0; g <= g1; g++) gs[g - g0] = gstart[g] - gstart[g0];
            rc = skb_sketch_batch(global_ctx(), &p, seed ? 1 : 0, g1 - g0, gs.data(), ptrs.data() + gstart[g0], lens.data() + gstart[g0],
                                  handles.data() + g0);
            g0 = g1;
        }
    }
    for (uint32_t g = 0; g < n; g++) if (handles[g]) { out[g].handle = std::make_shared<SketchHandle>(); out[g].handle->h = handles[g]; }
    check(global_ctx(), rc);
    return out;
}

// The reference resolves learned_ani=None to use_learned_ani(c, false, false, median) = (c >= 70 && !median) and then
// corrects the estimate with the model embedded in skani (lib.rs:611-614).  Without a model file this implementation
// returns the uncorrected estimate; it must not do so silently.
void warn_if_uncorrected(const Database& db, const py::object& learned_ani, bool median, bool robust) {
    static bool warned = false;
    if (warned || db.model || !learned_ani.is_none() || median || robust || db.params.c < 70) return;
    warned = true;
    if (PyErr_WarnEx(PyExc_RuntimeWarning,
                     "pyskani applies skani's learned ANI regression by default here (compression >= 70, no median); its model is "
                     "embedded in the skani crate and not available to pyskani_b200, so the UNCORRECTED estimate is returned. Load a "
                     "gbdt-rs JSON model (Database(model=...), Database.set_model(), $PYSKANI_B200_MODEL) or pass learned_ani=False.",
                     1) < 0)
        throw py::error_already_set();
}

void raise_query_error(int rc, const std::string& err) {
    switch (rc) {
        case SKB_OK: return;
        case SKB_ERR_ARG: throw py::value_error(err);
        case SKB_ERR_KEY: throw py::key_error(err);
        case SKB_ERR_NOMEM: throw std::bad_alloc();
        default: throw std::runtime_error(err);
    }
}

}  // namespace

PYBIND11_MODULE(_skani, m) {
    m.doc() = "A Python module for metagenomic sequence comparison with ``skani``, running on NVIDIA B200 GPUs.";
    m.attr("__package__") = "pyskani_b200";
    m.attr("__version__") = "0.1.0";
    m.attr("__author__") = "pyskani_b200 developers";
    py::dict build, deps;
    deps["skani"] = "0.3.0";   // version of the algorithm restated by the CUDA kernels (Cargo.toml:30-35 of the reference)
    deps["libskb"] = skb_version();
    build["dependencies"] = deps;
    build["target"] = "sm_100a";
    m.attr("__build__") = build;

    py::register_exception_translator([](std::exception_ptr p) {
        try { if (p) std::rethrow_exception(p); }
        catch (const OsError& e) {
            PyObject* cls = e.code == EEXIST ? PyExc_FileExistsError : PyExc_OSError;
            PyObject* args = Py_BuildValue("(is)", e.code, e.msg.c_str());
            PyErr_SetObject(cls, args);
            Py_XDECREF(args);
        }
    });

    // Test hooks of the on-disk layout (no device needed): the very writer / reader Database.save, flush, load and open
    // use, applied to host arrays.  tests/test_bincode_layout.py pins them byte for byte against an independent encoder.
    m.def("_encode_sketch", [](const std::string& name, uint64_t c, uint64_t k, uint64_t marker_c, bool has_seeds,
                               const std::vector<uint64_t>& kmer, const std::vector<uint32_t>& pos, const std::vector<uint32_t>& contig,
                               const std::vector<uint8_t>& canonical, const std::vector<std::string>& contigs, uint64_t total_len,
                               const std::vector<uint32_t>& contig_lengths, const std::vector<uint64_t>& markers, bool markers_only,
                               bool with_params) {
        HostSketch s;
        s.file_name = name; s.has_seeds = has_seeds; s.kmer = kmer; s.pos = pos; s.contig = contig; s.canonical = canonical;
        s.contigs = contigs; s.total_len = total_len; s.contig_lengths = contig_lengths; s.markers = markers;
        s.params = Params{c, k, marker_c};
        Writer w;
        if (with_params) write_params(w, s.params);
        write_sketch(w, s, markers_only);
        return py::bytes(w.buf);
    });
    m.def("_decode_sketch", [](const py::bytes& raw, bool with_params) {
        const std::string buf = raw;
        Reader r{(const uint8_t*)buf.data(), buf.size()};
        py::dict d;
        if (with_params) { Params p = read_params(r); d["params"] = py::make_tuple(p.c, p.k, p.marker_c); }
        HostSketch s = read_sketch(r);
        d["file_name"] = s.file_name; d["has_seeds"] = s.has_seeds; d["kmer"] = s.kmer; d["pos"] = s.pos; d["contig"] = s.contig;
        d["canonical"] = s.canonical; d["contigs"] = s.contigs; d["total_len"] = s.total_len; d["contig_lengths"] = s.contig_lengths;
        d["markers"] = s.markers; d["sketch_params"] = py::make_tuple(s.params.c, s.params.k, s.params.marker_c);
        d["consumed"] = r.off;
        return d;
    });
    m.def("_encode_index", [](const std::vector<std::tuple<std::string, uint64_t, uint64_t>>& entries) {
        Writer w;
        w.u64(entries.size());
        for (auto& e : entries) { w.str(std::get<0>(e)); w.u64(std::get<1>(e)); w.u64(std::get<2>(e)); }
        return py::bytes(w.buf);
    });

    py::class_<Hit>(m, "Hit", "A single hit found when querying a `~pyskani.Database` with a genome.")
        .def(py::init([](float identity, const std::string& query_name, float query_fraction, const std::string& reference_name,
                         float reference_fraction) {
                 auto fmt = [](float v) { return py::str(py::float_(v)).cast<std::string>(); };
                 if (identity < 0.0f || identity > 1.0f) throw py::value_error("Invalid value for `identity`: " + fmt(identity));
                 if (query_fraction < 0.0f || query_fraction > 1.0f) throw py::value_error("Invalid value for `query_fraction`: " + fmt(query_fraction));
                 if (reference_fraction < 0.0f || reference_fraction > 1.0f) throw py::value_error("Invalid value for `reference_fraction`: " + fmt(reference_fraction));
                 return Hit{identity, query_name, query_fraction, reference_name, reference_fraction};
             }),
             py::arg("identity"), py::arg("query_name"), py::arg("query_fraction"), py::arg("reference_name"), py::arg("reference_fraction"))
        .def("__repr__", [](const Hit& h) {
            return py::str("Hit(identity={!r}, query_name={!r}, query_fraction={!r}, reference_name={!r}, reference_fraction={!r})")
                .format(h.identity, h.query_name, h.query_fraction, h.reference_name, h.reference_fraction);
        })
        .def_property_readonly("identity", [](const Hit& h) { return h.identity; })
        .def_property_readonly("query_name", [](const Hit& h) { return h.query_name; })
        .def_property_readonly("query_fraction", [](const Hit& h) { return h.query_fraction; })
        .def_property_readonly("reference_name", [](const Hit& h) { return h.reference_name; })
        .def_property_readonly("reference_fraction", [](const Hit& h) { return h.reference_fraction; });

    py::class_<Sketch>(m, "Sketch", "A sketched genome.")
        .def_property_readonly("name", [](const Sketch& s) { return s.name; })
        .def_property_readonly("c", [](const Sketch& s) { return s.c; })
        .def_property_readonly("amino_acid", [](const Sketch& s) { return s.amino_acid; });

    py::class_<Database>(m, "Database", "A database storing sketched genomes.")
        .def(py::init([](const py::object& path, uint64_t compression, uint64_t marker_compression, uint64_t k, const py::object& format,
                         const py::object& model) {
                 auto db = std::make_unique<Database>();
                 db->model_from(model);
                 if (k < 1 || k > 16) throw py::value_error("Value of k > 16 for DNA; not allowed.");
                 if (compression < 1 || marker_compression < 1) throw py::value_error("compression factors must be positive");
                 db->params = Params{compression, k, marker_compression};
                 if (!path.is_none()) {
                     db->folder = fsdecode(path);
                     if (!exists(db->folder)) mkdirs(db->folder);
                     if (exists(join(db->folder, "markers.bin"))) throw_os(EEXIST, join(db->folder, "markers.bin"));   // lib.rs:395-399
                     db->storage = parse_format(format);
                 }
                 return db;
             }),
             py::arg("path") = py::none(), py::kw_only(), py::arg("compression") = 125, py::arg("marker_compression") = 1000,
             py::arg("k") = 15, py::arg("format") = py::none(), py::arg("model") = py::none())
        .def_static("load", [](const py::object& path, const py::object& model) { return open_impl(path, true, model); }, py::arg("path"),
                    py::kw_only(), py::arg("model") = py::none(),
                    "Load a database from a folder containing sketches (detached from the folder).")
        .def_static("open", [](const py::object& path, const py::object& model) { return open_impl(path, false, model); }, py::arg("path"),
                    py::kw_only(), py::arg("model") = py::none(),
                    "Open a database from a folder containing sketches; new sketches are appended to it.")
        .def("set_model", [](Database& db, const py::object& path) {
                 std::shared_ptr<ModelHandle> m = path.is_none() ? nullptr : load_model_file(fsdecode(path));
                 py::gil_scoped_release nogil;
                 std::unique_lock<std::shared_mutex> lk(db.mu);
                 db.set_model(std::move(m));
             }, py::arg("path"),
             "Load skani's learned-ANI regression (a gbdt-rs JSON dump) from a file; None removes it (extension over pyskani).")
        .def_property_readonly("has_model", [](const Database& db) { return (bool)db.model; })
        .def("__enter__", [](py::object self) { return self; })
        .def("__exit__", [](Database& db, const py::object&, const py::object&, const py::object&) {
            { py::gil_scoped_release nogil; db.flush(); }
            return false;
        })
        .def_property_readonly("path", [](const Database& db) -> py::object {
            if (db.storage == Storage::Memory) return py::none();
            return py::module_::import("pathlib").attr("Path")(db.folder);
        })
        .def_property_readonly("compression", [](const Database& db) { return db.params.c; })
        .def_property_readonly("marker_compression", [](const Database& db) { return db.params.marker_c; })
        .def("__len__", [](const Database& db) { return db.items.size(); })
        .def("sketch", [](Database& db, const std::string& name, const py::args& contigs, bool seed) {
                 Sketch s = sketch_impl(db, name, contigs, seed);
                 py::gil_scoped_release nogil;                       // before the lock, never the other way round
                 std::unique_lock<std::shared_mutex> lk(db.mu);
                 db.add(std::move(s), true);
             }, py::arg("name"), py::arg("seed") = true, "Add a reference genome to the database.")
        .def("query", [](Database& db, const std::string& name, const py::args& contigs, bool seed, const py::object& learned_ani,
                         bool median, bool robust, const py::object& cutoff, bool faster_small) {
                 warn_if_uncorrected(db, learned_ani, median, robust);
                 Sketch q = sketch_impl(db, name, contigs, seed);
                 skb_query_opts_t o{};
                 o.cutoff = cutoff.is_none() ? 0.0 : cutoff.cast<double>();
                 o.learned_ani = learned_ani.is_none() ? -1 : (learned_ani.cast<bool>() ? 1 : 0);
                 o.median = median; o.robust = robust; o.faster_small = faster_small;
                 skb_hit_t* hits = nullptr; uint64_t n = 0;
                 int rc;
                 std::string err;
                 std::vector<Hit> out;
                 {
                     // GIL first, lock second (lib.rs:569 allow_threads, then RwLock::read at lib.rs:617-621): concurrent
                     // query() calls on one Database share the lock; the GPU work itself is serialised inside libskb
                     py::gil_scoped_release nogil;
                     std::shared_lock<std::shared_mutex> lk(db.mu);
                     skb_sketch_t* qh = q.handle->h;
                     rc = skb_db_query(db.db, 1, &qh, &o, &hits, &n, nullptr);
                     if (rc != SKB_OK) err = skb_last_error(global_ctx());
                     else for (uint64_t i = 0; i < n; i++)
                         out.push_back(Hit{hits[i].ani, name, hits[i].af_query, db.items[hits[i].ref_index].name, hits[i].af_ref});
                     skb_hits_free(hits);
                 }
                 raise_query_error(rc, err);
                 return out;
             }, py::arg("name"), py::arg("seed") = true, py::arg("learned_ani") = py::none(), py::arg("median") = false,
             py::arg("robust") = false, py::arg("cutoff") = py::none(), py::arg("faster_small") = false,
             "Query the database with a genome.")
        .def("sketch_many", [](Database& db, const py::sequence& items, bool seed) {
                 std::vector<Sketch> sk = sketch_many_impl(db, items, seed);
                 py::gil_scoped_release nogil;
                 std::unique_lock<std::shared_mutex> lk(db.mu);
                 for (auto& s : sk) db.add(std::move(s), true);
             }, py::arg("items"), py::kw_only(), py::arg("seed") = true,
             "Add many reference genomes in one GPU batch: items = [(name, contigs), ...] (extension over pyskani).")
        .def("query_many", [](Database& db, const py::sequence& items, bool seed, const py::object& learned_ani, bool median, bool robust,
                              const py::object& cutoff, bool faster_small) {
                 warn_if_uncorrected(db, learned_ani, median, robust);
                 std::vector<Sketch> qs = sketch_many_impl(db, items, seed);
                 skb_query_opts_t o{};
                 o.cutoff = cutoff.is_none() ? 0.0 : cutoff.cast<double>();
                 o.learned_ani = learned_ani.is_none() ? -1 : (learned_ani.cast<bool>() ? 1 : 0);
                 o.median = median; o.robust = robust; o.faster_small = faster_small;
                 std::vector<skb_sketch_t*> qh;
                 for (auto& q : qs) qh.push_back(q.handle->h);
                 skb_hit_t* hits = nullptr; uint64_t n = 0;
                 int rc;
                 std::string err;
                 std::vector<std::vector<Hit>> out(qs.size());
                 {
                     py::gil_scoped_release nogil;
                     std::shared_lock<std::shared_mutex> lk(db.mu);
                     rc = skb_db_query(db.db, (uint32_t)qh.size(), qh.data(), &o, &hits, &n, nullptr);
                     if (rc != SKB_OK) err = skb_last_error(global_ctx());
                     else for (uint64_t i = 0; i < n; i++)
                         out[hits[i].query_index].push_back(Hit{hits[i].ani, qs[hits[i].query_index].name, hits[i].af_query,
                                                                db.items[hits[i].ref_index].name, hits[i].af_ref});
                     skb_hits_free(hits);
                 }
                 raise_query_error(rc, err);
                 return out;
             }, py::arg("items"), py::kw_only(), py::arg("seed") = true, py::arg("learned_ani") = py::none(), py::arg("median") = false,
             py::arg("robust") = false, py::arg("cutoff") = py::none(), py::arg("faster_small") = false,
             "Query with many genomes in one GPU batch: items = [(name, contigs), ...]; returns one list of Hit per query "
             "(extension over pyskani).")
        .def("save", [](Database& db, const py::object& path, bool overwrite, const py::object& format, bool strict_format) {
                 const std::string folder = fsdecode(path);
                 if (!exists(folder)) mkdirs(folder);
                 const std::string markers = join(folder, "markers.bin");
                 if (!overwrite && exists(markers)) throw_os(EEXIST, markers);
                 // The reference maps the two format names the other way round in save() (lib.rs:696-699): None and
                 // "consolidated" write one `<name>.sketch` per genome, "separated" writes sketches.db + index.db.  That is
                 // what existing pyskani callers get on disk, so it is the default here too; strict_format=True writes the
                 // layout the name says.  Database.load / Database.open read either layout.
                 Storage st = parse_format(format);
                 if (!strict_format) st = st == Storage::Consolidated ? Storage::Folder : Storage::Consolidated;
                 py::gil_scoped_release nogil;
                 std::shared_lock<std::shared_mutex> lk(db.mu);
                 db.save_markers(markers);
                 std::vector<IndexEntry> idx;
                 for (auto& s : db.items) {
                     Writer w;
                     write_params(w, db.params);
                     write_sketch(w, export_sketch(s, db.params), false);
                     // DatabaseStorage::store (lib.rs:56-62 / 64-88): file named after Sketch.file_name; sketches.db is opened
                     // for appending, so an existing one grows
                     if (st == Storage::Folder) write_file(join(folder, s.name + ".sketch"), w.buf, "wb");
                     else { idx.push_back(IndexEntry{s.name, file_size(join(folder, "sketches.db")), (uint64_t)w.buf.size()}); write_file(join(folder, "sketches.db"), w.buf, "ab"); }
                 }
                 if (st == Storage::Consolidated) Database::save_index(join(folder, "index.db"), idx);
             }, py::arg("path"), py::arg("overwrite") = false, py::arg("format") = py::none(), py::kw_only(), py::arg("strict_format") = false,
             "Save the database to the given path (format names mapped as the reference maps them; strict_format=True "
             "writes the layout the name says).")
        .def("flush", [](Database& db) { py::gil_scoped_release nogil; db.flush(); }, "Flush the database buffers to disk.");
}
