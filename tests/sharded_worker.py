"""Worker of tests/test_sharded.py::test_group_torchrun_processes_match_single_gpu (launched through torch.distributed.run).

Every rank is a process with its own libkcgpu context; the ranks' heaps are mapped into each other through CUDA IPC handles
(gathered over gloo: CPU plumbing only, so that the test also runs with all ranks on ONE GPU, which NCCL would refuse).  Rank 0
compares the concatenated slices with kc_compute on a single GPU and writes the verdict."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kmercamel_b200 as kb  # noqa: E402
from kmercamel_b200 import sharded, synth  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dev = int(os.environ.get("LOCAL_RANK", "0")) % torch.cuda.device_count()
    torch.cuda.set_device(dev)
    ctx = kb.Context(dev, torch.cuda.current_stream().cuda_stream)
    g = synth.frame_records(synth.random_genome_records(5, 600_000, 41))[0]
    rep = synth.frame_records(synth.human_like_genome(2_000_000, 43))[0]
    reads = synth.frame_reads(synth.reads_chunks(60_000, 12.0, 150, 0.01, 44, chunk_reads=2000), 150)
    cases = [(g, 31, True, 1), (rep, 31, True, 1), (np.concatenate([g, g[:len(g) // 2]]), 31, False, 1), (g, 63, True, 1), (reads, 31, True, 2),
             (rep, 127, False, 1), (g, 31, True, 1)]
    cap = max(len(c[0]) for c in cases)
    sharded.attach(ctx, rank, world, k=127, n_bytes_cap=cap)
    ok, n = True, 0
    single = kb.Context(dev) if rank == 0 else None
    for seq, k, compl, z in cases:
        d = torch.from_numpy(seq).cuda()
        torch.cuda.synchronize()
        r = sharded.sharded_compute(ctx, d.data_ptr(), d.numel(), k=k, complements=compl, min_frequency=z)
        part = np.frombuffer(ctx.copy_to_host(r.ms_ptr, r.slice_len), dtype=np.uint8)
        meta = torch.tensor([r.slice_begin, r.slice_len, r.length, r.n_kmers], dtype=torch.int64)
        metas = [torch.empty_like(meta) for _ in range(world)]
        dist.all_gather(metas, meta)
        buf = torch.zeros(int(r.length) + 64, dtype=torch.uint8)
        buf[r.slice_begin:r.slice_begin + r.slice_len] = torch.from_numpy(part.copy())
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)             # disjoint slices: SUM == concatenation
        if rank == 0:
            want = single.compute(seq, k=k, complements=compl, min_frequency=z)
            at = 0
            for m in metas:
                ok &= int(m[0]) == at and int(m[2]) == want.length and int(m[3]) == want.n_kmers
                at += int(m[1])
            ok &= at == want.length and buf[:want.length].numpy().tobytes() == want.ms
        n += 1
        dist.barrier()                                          # `d` is freed below: no rank may still be reading ITS copy... (own copy only)
    fast = ctx.stat("fast_runs")
    if rank == 0:
        json.dump({"world": world, "cases": n, "all_identical": bool(ok), "fast_runs": fast, "devices": torch.cuda.device_count()}, open(sys.argv[1], "w"))
    ctx.group_close()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
