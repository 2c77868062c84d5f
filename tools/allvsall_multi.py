"""All-vs-all over N GPUs with pyskani_b200.parallel (NCCL exchange of the sketch database).
torchrun --nproc-per-node N tools/allvsall_multi.py [families] [genome_len]; rank 0 checks against a 1-GPU run."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from pyskani_b200 import parallel, synth

F = int(sys.argv[1]) if len(sys.argv) > 1 else 8
glen = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
rank = int(os.environ.get("RANK", 0)); local = int(os.environ.get("LOCAL_RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
genomes = []
for f in range(F):
    base = synth.random_genome(glen, 3000 + f)
    genomes += [[base.tobytes()]] + [[synth.mutate(base, d, 4000 + 10 * f + i).tobytes()] for i, d in enumerate((0.02, 0.06, 0.12))]
be = parallel.CudaBackend(local)
for it in range(2):
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    table = parallel.all_vs_all(genomes, be, dist=dist if world > 1 else None, device=torch.device("cuda", local))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
if rank == 0:
    n = len(genomes)
    print(f"world {world}: {n} genomes, {n*n} ordered pairs, {len(table)} hits, {1e3*dt:.1f} ms")
    single = parallel.all_vs_all(genomes, be) if world > 1 else table
    same = table.shape == single.shape and np.array_equal(table, single)
    print("identical to the single-GPU table:", same)
    assert same
if world > 1:
    dist.barrier(); dist.destroy_process_group()
