"""Host-buffer sketching with contigs laid out back to back in one buffer (copies merged) against the same contigs as
separate objects (one copy each): the sketches must be identical."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyskani_b200 import capi, synth

lens = [300_000, 262_144, 1_000_003, 700, 400_000, 262_145, 5_000_000, 263_000]
seqs = [synth.random_genome(l, 500 + i) for i, l in enumerate(lens)]
big = np.zeros(sum((l + 15) // 16 * 16 + 16 for l in lens) + 64, np.uint8)
views, off = [], 16
for s in seqs:
    big[off:off + len(s)] = s
    views.append(big[off:off + len(s)])
    off += (len(s) + 15) // 16 * 16 + 16
ctx = capi.Context(0)
genomes_a = [[views[0], views[1], views[2]], [views[3], views[4]], [views[5], views[6], views[7]]]
genomes_b = [[seqs[0].tobytes(), seqs[1].tobytes(), seqs[2].tobytes()], [seqs[3].tobytes(), seqs[4].tobytes()],
             [seqs[5].tobytes(), seqs[6].tobytes(), seqs[7].tobytes()]]
a = ctx.sketch_batch(genomes_a)
b = ctx.sketch_batch(genomes_b)
for x, y in zip(a, b):
    ex, ey = x.export(), y.export()
    for k in ex:
        assert np.array_equal(ex[k], ey[k]), k
print("merged and separate copies give identical sketches:", [s.info().n_seeds for s in a])
