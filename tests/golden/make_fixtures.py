"""Generate tests/golden/ecoli_pair.npz from the reference's own test fixtures.

Run in the build container only (the reference tree does not exist on the GPU box):
    python tests/golden/make_fixtures.py

Source: /root/reference/src/pyskani/tests/e.coli-EC590.fasta.gz and e.coli-K12.fasta.gz, the two
genomes the reference's known-answer test sketches and queries (test_ani.py:22-26).  Both records are
single-contig, upper-case ACGT only, so a 2-bit packing (A=0 C=1 G=2 T=3, 4 bases/byte, first base in
the low bits) is lossless.  The golden ANI/AF values asserted by the reference (test_ani.py:28-61) are
stored next to the sequences.
"""
import gzip, hashlib, os, sys
import numpy as np

REF_TESTS = "/root/reference/src/pyskani/tests"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ecoli_pair.npz")


def read_single_fasta(path):
    name, seq = None, []
    with gzip.open(path, "rt") as f:
        for line in f:
            if line.startswith(">"):
                if name is not None:
                    raise SystemExit("expected one record in " + path)
                name = line[1:].split()[0]
            else:
                seq.append(line.strip())
    return name, "".join(seq).encode("ascii")


def pack2(seq: bytes):
    a = np.frombuffer(seq, dtype=np.uint8)
    lut = np.full(256, 255, np.uint8)
    for i, ch in enumerate(b"ACGT"):
        lut[ch] = i
    codes = lut[a]
    if (codes == 255).any():
        raise SystemExit("non-ACGT byte in fixture: 2-bit packing would be lossy")
    pad = (-len(codes)) % 4
    codes = np.concatenate([codes, np.zeros(pad, np.uint8)]).reshape(-1, 4)
    return (codes[:, 0] | (codes[:, 1] << 2) | (codes[:, 2] << 4) | (codes[:, 3] << 6)).astype(np.uint8)


def main():
    out = {}
    for key, fn in (("EC590", "e.coli-EC590.fasta.gz"), ("K12", "e.coli-K12.fasta.gz")):
        name, seq = read_single_fasta(os.path.join(REF_TESTS, fn))
        out[key + "_packed"] = pack2(seq)
        out[key + "_len"] = np.int64(len(seq))
        out[key + "_sha256"] = np.bytes_(hashlib.sha256(seq).hexdigest())
        out[key + "_record"] = np.bytes_(name)
        print(key, name, len(seq), hashlib.sha256(seq).hexdigest()[:16])
    # reference/src/pyskani/tests/test_ani.py:28-61 (assertAlmostEqual places=4)
    out["golden_af_ref"] = np.float64(0.9246)
    out["golden_af_query"] = np.float64(0.9189)
    out["golden_ani_learned"] = np.float64(0.9939)
    out["golden_ani_no_learned"] = np.float64(0.9946)
    out["golden_ani_robust"] = np.float64(0.9977)
    out["golden_ani_median"] = np.float64(0.9995)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT))


if __name__ == "__main__":
    sys.exit(main())
