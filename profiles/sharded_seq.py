import hashlib, os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch, torch.distributed as dist
import kmercamel_b200 as kb
from kmercamel_b200 import sharded, synth
rank = int(os.environ.get("RANK", 0)); lr = int(os.environ.get("LOCAL_RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
dev = torch.device("cuda", lr)
K = 31
ctx = kb.Context(lr, torch.cuda.current_stream().cuda_stream)
comm = sharded.TorchComm(dev)
first = True
for name, seed in (("A", 12345 + rank), ("B", 777), ("B", 777), ("A", 12345 + rank), ("A", 12345 + rank), ("B", 777)):
    part, _, _ = synth.frame_records(synth.random_genome_records(50, 1_000_000, seed))
    own = torch.from_numpy(part).to(dev)
    full = torch.empty(world * own.numel(), dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(full, own)
    torch.cuda.synchronize()
    ops = sharded.GpuOps(ctx, full)
    if first:
        ops.setup_p2p(comm, K); first = False
    r = sharded.sharded_compute_p2p(ops, comm, full.numel(), k=K, min_frequency=1)
    print(f"rank {rank} {name}: kept={r.n_kept} occ={r.n_occurrences} sent={r.items_sent} recv={r.items_received}" + (f" len={r.result.length}" if r.result else ""), flush=True)
    dist.barrier()
dist.destroy_process_group()
