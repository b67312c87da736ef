"""One rank of the world_size-2 CPU test (gloo): aligns its length-balanced shard of the golden reads
through libgcalign built against the C-ABI test double and sends the decoded records to rank 0,
which compares the union with the reference's golden GAM."""
import os
import sys

import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from graphchainer_b200 import align, gam, shard  # noqa: E402


def main():
    idx_path, fasta, golden_gam, out_flag, lib_path = sys.argv[1:6]  # lib_path: libgcalign built against the C-ABI test double
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    batch = align.ReadBatch.from_fasta(fasta)
    mine = shard.length_balanced_shards(batch.lengths(), world)[rank]
    sub = batch.subset(mine)
    aligner = align.Aligner(idx_path, device=0, host_threads=2, streams=2, lib_path=lib_path)
    blob, summ, st = aligner.align(sub, gam=True)
    aligner.close()
    local = {}
    for k, i in enumerate(mine):
        off, size = int(summ[k]["gam_offset"]), int(summ[k]["gam_size"])
        if size:
            local[int(i)] = bytes(blob[off:off + size])
    merged = shard.gather_records(local, rank, world)
    total_bp = sum(int(x) for x in batch.lengths())
    shard_bp = [None] * world
    dist.all_gather_object(shard_bp, int(sub.total_bp))
    if rank == 0:
        tmp = out_flag + ".gam"
        with open(tmp, "wb") as f:
            for i in sorted(merged):
                f.write(merged[i])
        diffs = gam.diff_gam(gam.read_gam(tmp), gam.read_gam(golden_gam))
        balanced = max(shard_bp) - min(shard_bp) <= max(int(x) for x in batch.lengths())
        ok = not diffs and sum(shard_bp) == total_bp and balanced
        with open(out_flag, "w") as f:
            f.write("OK" if ok else f"FAIL diffs={diffs[:3]} shard_bp={shard_bp}")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
