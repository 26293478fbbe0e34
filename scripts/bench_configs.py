"""The other BASELINE.json configs as measured parity cases (not the contract bench line):
  config 2: NTT / INTT microbenchmark over the key-level limbs, batch 1..65,536 polynomials
  config 4: mixed precompile batch (add/sub/mul x ct.ct/ct.pt/pt.ct x 4 types) device-resident, cost-weighted
  config 5: pk-encrypt of 16,384 plaintexts under the network key + decrypt round trip
Writes one JSON object to stdout. Run on a B200: python scripts/bench_configs.py"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fhe_precompiles_b200 import device as fdev  # noqa: E402

N = 4096
MODULI = (0xFFFFEE001, 0xFFFFC4001, 0x1FFFFE0001)


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def rand_ct(n, seed):
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    out = torch.empty((n, 2, 2, N), dtype=torch.int64, device="cuda")
    for l in range(2):
        out[:, :, l, :] = torch.randint(0, MODULI[l], (n, 2, N), generator=g, device="cuda", dtype=torch.int64)
    return out


def run_configs(rank: int, world: int, dist, quick: bool = False) -> dict:
    """Configs 2, 4 and 5 on the current CUDA device of every rank (fdev.init done by the caller); returns the whole-job record
    (identical on every rank up to rank-0-only fields).  quick: fewer NTT batch sizes and repetitions (bench.py's `configs`)."""
    from fhe_precompiles_b200.sharding import call_cost, max_over_ranks, shard_by_cost, shard_range

    def sync():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    iters = 2 if quick else 3
    res = {"n_gpus": world}
    # ---- config 2 (one GPU by definition: every rank runs it, rank 0's numbers are reported)
    ntt = []
    g = torch.Generator(device="cuda")
    g.manual_seed(1)
    batches = (1, 256, 4096, 65536) if quick else (1, 4, 16, 64, 256, 1024, 4096, 16384, 65536)
    for batch in batches:
        x = torch.empty((batch, 3, N), dtype=torch.int64, device="cuda")
        for l in range(3):
            x[:, l, :] = torch.randint(0, MODULI[l], (batch, N), generator=g, device="cuda", dtype=torch.int64)
        f = timeit(lambda: fdev.ntt_(x, [0, 1, 2]), iters=iters + 2)
        i = timeit(lambda: fdev.ntt_(x, [0, 1, 2], inverse=True), iters=iters + 2)
        ntt.append({"polys": batch, "limbs": 3 * batch, "fwd_ms": f, "inv_ms": i,
                    "fwd_Mlimb_per_s": 3 * batch / f / 1e3, "inv_Mlimb_per_s": 3 * batch / i / 1e3,
                    "fwd_GBps": 3 * batch * 65536 / f / 1e6})
        del x
    res["config2_ntt"] = ntt
    # ---- config 4: 65,536 mixed calls (device-resident operands), grouped by kernel family
    net_pub = open(os.path.join(ROOT, "fhe_precompiles_b200/data/network.pub"), "rb").read()
    pk_h, rk_h = fdev.parse_public_key(net_pub)
    rk, pk = rk_h.cuda(), pk_h.cuda()
    rng = np.random.default_rng(3)
    n_calls = 65536
    ops = rng.integers(0, 3, n_calls)      # add, sub, mul
    shapes = rng.integers(0, 3, n_calls)   # ctct, ctpt, ptct
    # cost-weighted sharding of the calls over the ranks (mul ct x ct dominates); every rank runs only its share
    names = [f"{('add', 'sub', 'mul')[o]}_" + ("cipheri64_cipheri64" if s == 0 else "cipheri64_i64" if s == 1 else "i64_cipheri64")
             for o, s in zip(ops, shapes)]
    mine = shard_by_cost([call_cost(nm) for nm in names], world)[rank]
    ops, shapes = ops[mine], shapes[mine]
    counts = {(o, s): int(((ops == o) & (shapes == s)).sum()) for o in range(3) for s in range(3)}
    pool = 4096
    a, b = rand_ct(pool, 10), rand_ct(pool, 11)
    plain = torch.randint(0, 2, (pool, N), device="cuda", dtype=torch.int16)
    out = torch.empty_like(a)

    def run_family(o, s, n):
        done = 0
        while done < n:
            c = min(pool, n - done)
            if s == 0:
                if o == 0:
                    fdev.add(a[:c], b[:c])
                elif o == 1:
                    fdev.sub(a[:c], b[:c])
                else:
                    fdev.mul_relin(a[:c], b[:c], rk, out=out[:c])
            else:
                if o == 2:
                    fdev.multiply_plain(a[:c], plain[:c])
                else:
                    fdev.plain_addsub(a[:c], plain[:c], 0 if o == 0 else (1 if s == 1 else 3))
            done += c

    def run_all():
        for (o, s), n in counts.items():
            run_family(o, s, n)

    sync()
    ms = max_over_ranks(timeit(run_all, iters=iters, warm=1), dist)
    sync()
    res["config4_mixed"] = {"calls": n_calls, "ms": ms, "calls_per_s": n_calls / ms * 1e3, "calls_on_rank0": len(mine),
                            "mix": {f"{'add sub mul'.split()[o]}_{'ctct ctpt ptct'.split()[s]}": n for (o, s), n in counts.items()}}
    del a, b, plain, out
    # ---- config 5: encrypt 16,384 random i64 under the network key, decrypt, compare
    n_total = 16384
    lo, hi = shard_range(n_total, rank, world)
    n = hi - lo
    vals = rng.integers(-(2**62), 2**62, n_total)[lo:hi]
    mag = np.abs(vals).astype(np.uint64)
    bits = ((mag[:, None] >> np.arange(64, dtype=np.uint64)[None, :]) & 1).astype(np.int64)
    coeff = np.where(vals[:, None] < 0, bits * 4095, bits)
    pl = np.zeros((n, N), dtype=np.int16)
    pl[:, :64] = coeff.astype(np.uint16).view(np.int16)
    dpl = torch.from_numpy(pl).cuda()
    seeds = torch.from_numpy(np.random.default_rng(7 + rank).integers(0, 2**63, size=(n, 8), dtype=np.int64)).cuda()  # 512 bits per op
    sk = torch.empty((3, N), dtype=torch.int64)
    from fhe_precompiles_b200 import _lib
    pri = open(os.path.join(ROOT, "fhe_precompiles_b200/data/network.pri"), "rb").read()
    assert _lib.lib().fhe_b200_parse_private_key(pri, len(pri), sk.data_ptr()) == 0
    dsk = sk.cuda()
    sync()
    enc_ms = max_over_ranks(timeit(lambda: fdev.encrypt(pk, dpl, seeds), iters=iters, warm=1), dist)
    ct = fdev.encrypt(pk, dpl, seeds)
    sync()
    dec_ms = max_over_ranks(timeit(lambda: fdev.decrypt_checked(ct, dsk), iters=iters, warm=1), dist)
    back, flags = fdev.decrypt_checked(ct, dsk)
    ok = float(torch.equal(back, dpl) and not bool(flags.any()))
    ok = -max_over_ranks(-ok, dist)  # min over ranks
    res["config5_encrypt_decrypt"] = {"plaintexts": n_total, "encrypt_ms": enc_ms, "encrypt_per_s": n_total / enc_ms * 1e3,
                                      "decrypt_ms": dec_ms, "decrypt_per_s": n_total / dec_ms * 1e3,
                                      "roundtrip_all_equal": bool(ok)}
    return res


def main():
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist  # barrier + max-over-ranks only (gloo): the configs shard with no collective

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("gloo")
    fdev.init(local)
    res = run_configs(rank, world, dist)
    if rank == 0:
        print(json.dumps(res))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
