"""BASELINE.json configs[2..4] at (or near) full size on one B200: timing + size-independent properties.
Inputs are generated on the GPU with torch (seeded torch generators: same distributions as SURVEY.md §8d, not the numpy
streams).  usage: python profiles/run_big_configs.py [cfg3a cfg3b cfg4 cfg5 ...] [--scale F]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import kmercamel_b200 as kb

dev = torch.device("cuda", 0)
LUT = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
scale = 1.0
names = [a for a in sys.argv[1:] if not a.startswith("--")] or ["cfg3a", "cfg3b", "cfg4", "cfg5"]
for i, a in enumerate(sys.argv):
    if a == "--scale":
        scale = float(sys.argv[i + 1]); names = [n for n in names if n != sys.argv[i + 1]]


def genome(n_records, record_len, seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    out = torch.empty((n_records, record_len + 1), dtype=torch.uint8, device=dev)
    for r in range(n_records):
        out[r, :record_len] = LUT[torch.randint(0, 4, (record_len,), generator=g, device=dev, dtype=torch.uint8).long()]
    out[:, record_len] = 10
    return out.flatten()


def reads(genome_len, coverage, read_len, err, seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    codes = torch.randint(0, 4, (genome_len,), generator=g, device=dev, dtype=torch.uint8)
    n_reads = int(genome_len * coverage / read_len)
    out = torch.empty((n_reads, read_len + 1), dtype=torch.uint8, device=dev)
    ar = torch.arange(read_len, device=dev)
    for lo in range(0, n_reads, 1 << 20):
        hi = min(n_reads, lo + (1 << 20))
        starts = torch.randint(0, genome_len - read_len + 1, (hi - lo,), generator=g, device=dev)
        rd = codes[starts[:, None] + ar[None, :]]
        flip = torch.rand(hi - lo, generator=g, device=dev) < 0.5
        rd[flip] = (3 - rd[flip]).flip(1)
        e = torch.rand(rd.shape, generator=g, device=dev) < err
        rd[e] = (rd[e] + torch.randint(1, 4, (int(e.sum()),), generator=g, device=dev, dtype=torch.uint8)) % 4
        out[lo:hi, :read_len] = LUT[rd.long()]
    out[:, read_len] = 10
    return out.flatten()


def ms_to_spss(ms: torch.Tensor, k: int) -> torch.Tensor:
    """reference conversions.h:45-72 ms2spss on the GPU: every maximal run of ON letters + the following k-1 letters."""
    on = ms <= 90
    prev = torch.cat([torch.zeros(1, dtype=torch.bool, device=dev), on[:-1]])
    nxt = torch.cat([on[1:], torch.zeros(1, dtype=torch.bool, device=dev)])
    s = torch.nonzero(on & ~prev).flatten()
    e = torch.nonzero(on & ~nxt).flatten()
    ln = e - s + 1 + (k - 1) + 1                      # + '\n'
    off = torch.cumsum(ln, 0) - ln
    total = int(ln.sum())
    rec = torch.repeat_interleave(torch.arange(len(s), device=dev), ln)
    j = torch.arange(total, device=dev) - off[rec]
    src = (s[rec] + j).clamp(max=ms.numel() - 1)
    out = ms[src] & 0xDF                              # upper case
    out[j == ln[rec] - 1] = 10
    return out


def run(name, seq, k, complements, z, verify_set):
    ctx = kb.Context(0, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    res = {"config": name, "n_bytes": int(seq.numel()), "k": k, "complements": complements, "min_frequency": z}
    t0 = time.perf_counter(); r = ctx.compute_device(seq.data_ptr(), seq.numel(), k=k, complements=complements, min_frequency=z); t1 = time.perf_counter()
    r = ctx.compute_device(seq.data_ptr(), seq.numel(), k=k, complements=complements, min_frequency=z); t2 = time.perf_counter()
    res.update(first_call_s=t1 - t0, second_call_s=t2 - t1, device_ms=r.times_ms, n_kmers=r.n_kmers, n_occurrences=r.n_occurrences, nodes=r.n_nodes,
               length=r.length, launches=r.n_launches, kmers_per_s=r.n_kmers / (r.times_ms["total"] / 1e3), occurrences_per_s=r.n_occurrences / (r.times_ms["total"] / 1e3))
    ms_host = torch.frombuffer(bytearray(ctx.copy_to_host(r.ms_ptr, r.length)), dtype=torch.uint8)
    ctx.close()                      # give the arena back before the verification allocates
    ms = ms_host.to(dev)
    ones = int((ms <= 90).sum())
    res["ones_equal_kmers"] = ones == r.n_kmers
    res["tail_lower"] = bool((ms[-(k - 1):] > 90).all()) if k > 1 else True
    if verify_set:
        ctx2 = kb.Context(0)
        keys_in, cnt_in = ctx2.count_kmers(seq.cpu().numpy(), k=k, complements=complements, min_frequency=z)
        spss = ms_to_spss(ms, k)
        keys_ms, _ = ctx2.count_kmers(spss.cpu().numpy(), k=k, complements=complements)
        res["count_kmers_n"] = int(len(keys_in))
        res["set_equal"] = bool(keys_in.shape == keys_ms.shape and np.array_equal(keys_in, keys_ms))
        res["sorted"] = bool(np.all((keys_in[1:, -1] > keys_in[:-1, -1]) | (keys_in[1:, -1] == keys_in[:-1, -1])))
        ctx2.close()
    print(json.dumps(res), flush=True)
    return res


out = []
for n in names:
    torch.cuda.empty_cache()
    if n == "cfg3a":
        out.append(run("configs[2] 500 Mbp genome k=63 -u (u128)", genome(50, int(10_000_000 * scale), 31337), 63, False, 1, True))
    elif n == "cfg3b":
        out.append(run("configs[2] 500 Mbp genome k=127 -u (u256)", genome(50, int(10_000_000 * scale), 31337), 127, False, 1, True))
    elif n == "cfg4":
        out.append(run("configs[3] 30x reads of a 100 Mbp genome, k=31 -z 2", reads(int(100_000_000 * scale), 30.0, 150, 0.01, 2024), 31, True, 2, True))
    elif n == "cfg5":
        out.append(run("configs[4] 3.1 Gbp uniform genome k=31", genome(31, int(100_000_000 * scale), 3100), 31, True, 1, False))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "big_configs.json"), "w"), indent=1)
