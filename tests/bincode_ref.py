"""Independent bincode 1.3.3 encoder / decoder of pyskani's four database files (test infrastructure).

Written from SURVEY.md Appendix C, not from the product's C++ writer: bincode default options are little-endian,
fixed-width integers, usize -> u64, bool -> u8, Option -> u8 tag + value, String / Vec / HashMap / HashSet -> u64 length
prefix + elements, struct / tuple = fields in declaration order without names.

  <name>.sketch, blobs of sketches.db : (SketchParams, Sketch)                       reference lib.rs:57-62, 74-87
  markers.bin                         : (SketchParams, Vec<Sketch with kmer_seeds_k = None>)    lib.rs:187-197
  index.db                            : Vec<IndexEntry { file_name: String, offset: u64, length: u64 }> sorted by offset
                                                                                                lib.rs:203-211
  SketchParams { c, k, marker_c: usize, use_syncs, use_aa: bool, acgt_to_aa_encoding: Vec<u64>, acgt_to_aa_letters: Vec<u8>,
                 orf_size: usize }
  Sketch { file_name: String, kmer_seeds_k: Option<HashMap<u64, SmallVec<[SeedPosition; 1]>>>, contigs: Vec<String>,
           total_sequence_length: usize, contig_lengths: Vec<u32>, repetitive_kmers: usize, marker_seeds: HashSet<u64>,
           marker_c, c, k: usize, contig_order: usize, amino_acid: bool }
  SeedPosition { pos: u32, canonical: bool, contig_index: u32, phase: u8 }
Hash containers have no defined order: callers choose one when encoding and compare decoded content as sets.
"""
import struct

# standard genetic code, codons ordered by 2-bit bases A C G T (AAA, AAC, AAG, AAT, ACA, ...)
CODON_TABLE = "KNKNTTTTRSRSIIMIQHQHPPPPRRRRLLLLEDEDAAAAGGGGVVVV*Y*YSSSS*CWCLFLF"
AA_ORDER = "ACDEFGHIKLMNPQRSTVWY*"
ORF_SIZE = 30


def u8(v): return struct.pack("<B", v)
def u32(v): return struct.pack("<I", v)
def u64(v): return struct.pack("<Q", v)
def string(s): b = s.encode(); return u64(len(b)) + b


def encode_params(c, k, marker_c):
    out = u64(c) + u64(k) + u64(marker_c) + u8(0) + u8(0)
    out += u64(64) + b"".join(u64(AA_ORDER.index(a)) for a in CODON_TABLE)
    out += u64(64) + CODON_TABLE.encode()
    return out + u64(ORF_SIZE)


def encode_sketch(name, seeds, contigs, total_len, contig_lengths, markers, c, k, marker_c, seed_order=None, marker_order=None):
    """seeds: None (markers-only sketch) or dict kmer -> list of (pos, canonical, contig_index); orders: iteration order of
    the two hash containers (defaults: ascending)."""
    out = string(name)
    if seeds is None:
        out += u8(0)
    else:
        keys = list(seed_order) if seed_order is not None else sorted(seeds)
        out += u8(1) + u64(len(keys))
        for km in keys:
            out += u64(km) + u64(len(seeds[km]))
            for pos, canon, ci in seeds[km]:
                out += u32(pos) + u8(1 if canon else 0) + u32(ci) + u8(0)
    out += u64(len(contigs)) + b"".join(string(x) for x in contigs)
    out += u64(total_len)
    out += u64(len(contig_lengths)) + b"".join(u32(x) for x in contig_lengths)
    out += u64(0)
    ms = list(marker_order) if marker_order is not None else sorted(markers)
    out += u64(len(ms)) + b"".join(u64(x) for x in ms)
    out += u64(marker_c) + u64(c) + u64(k) + u64(0) + u8(0)
    return out


def encode_index(entries):
    entries = sorted(entries, key=lambda e: e[1])
    return u64(len(entries)) + b"".join(string(n) + u64(o) + u64(l) for n, o, l in entries)


class Cursor:
    def __init__(self, buf, off=0):
        self.buf, self.off = buf, off

    def take(self, fmt):
        v = struct.unpack_from(fmt, self.buf, self.off)
        self.off += struct.calcsize(fmt)
        return v[0]

    def string(self):
        n = self.take("<Q")
        s = self.buf[self.off:self.off + n].decode()
        assert len(s.encode()) == n, "truncated string"
        self.off += n
        return s


def decode_params(cur):
    c, k, mc = cur.take("<Q"), cur.take("<Q"), cur.take("<Q")
    syncs, aa = cur.take("<B"), cur.take("<B")
    enc = [cur.take("<Q") for _ in range(cur.take("<Q"))]
    letters = bytes(cur.take("<B") for _ in range(cur.take("<Q")))
    orf = cur.take("<Q")
    return dict(c=c, k=k, marker_c=mc, use_syncs=syncs, use_aa=aa, encoding=enc, letters=letters, orf_size=orf)


def decode_sketch(cur):
    d = dict(file_name=cur.string())
    tag = cur.take("<B")
    assert tag in (0, 1)
    d["seeds"] = None
    if tag:
        d["seeds"], d["seed_order"] = {}, []
        for _ in range(cur.take("<Q")):
            km, n = cur.take("<Q"), cur.take("<Q")
            lst = []
            for _ in range(n):
                pos, canon, ci, phase = cur.take("<I"), cur.take("<B"), cur.take("<I"), cur.take("<B")
                assert canon in (0, 1) and phase == 0
                lst.append((pos, bool(canon), ci))
            assert km not in d["seeds"]
            d["seeds"][km] = lst
            d["seed_order"].append(km)
    d["contigs"] = [cur.string() for _ in range(cur.take("<Q"))]
    d["total_len"] = cur.take("<Q")
    d["contig_lengths"] = [cur.take("<I") for _ in range(cur.take("<Q"))]
    d["repetitive_kmers"] = cur.take("<Q")
    d["markers"] = [cur.take("<Q") for _ in range(cur.take("<Q"))]
    d["marker_c"], d["c"], d["k"] = cur.take("<Q"), cur.take("<Q"), cur.take("<Q")
    d["contig_order"], d["amino_acid"] = cur.take("<Q"), cur.take("<B")
    return d


def decode_sketch_file(buf):
    cur = Cursor(buf)
    p = decode_params(cur)
    s = decode_sketch(cur)
    assert cur.off == len(buf), "trailing bytes"
    return p, s


def decode_markers_file(buf):
    cur = Cursor(buf)
    p = decode_params(cur)
    sk = [decode_sketch(cur) for _ in range(cur.take("<Q"))]
    assert cur.off == len(buf), "trailing bytes"
    return p, sk


def decode_index_file(buf):
    cur = Cursor(buf)
    out = [(cur.string(), cur.take("<Q"), cur.take("<Q")) for _ in range(cur.take("<Q"))]
    assert cur.off == len(buf)
    return out
