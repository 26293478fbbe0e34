"""Differential soak of the ct x ct multiply + relinearise path: many batches of random and adversarial ciphertext pairs through
the device-resident C-ABI entry point, bit-compared with the CPU oracle (all host threads).  Needs a GPU.
usage: python scripts/soak.py [batches] [ops_per_batch] [seed_base]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fhe_precompiles_b200 import device as fdev  # noqa: E402
from oracle import bfv  # noqa: E402  (checker)

N = 4096
Q = (0xFFFFEE001, 0xFFFFC4001)


def adversarial(rng, n):
    """extreme residues: 0, 1, q-1, q-2, values around 2^32 and 2^35, sparse and constant polynomials"""
    out = np.zeros((n, 2, 2, N), dtype=np.uint64)
    for i in range(n):
        for p in range(2):
            for l in range(2):
                q = Q[l]
                pool = np.array([0, 1, 2, q - 1, q - 2, q // 2, q // 2 + 1, (1 << 32) - 1, 1 << 32, (1 << 32) + 1, (1 << 35) - 1, 1 << 35],
                                dtype=np.uint64)
                kind = rng.integers(0, 4)
                if kind == 0:
                    out[i, p, l] = rng.choice(pool, size=N)
                elif kind == 1:
                    out[i, p, l] = pool[rng.integers(0, len(pool))]
                elif kind == 2:
                    v = np.zeros(N, dtype=np.uint64)
                    idx = rng.integers(0, N, size=8)
                    v[idx] = rng.choice(pool, size=8)
                    out[i, p, l] = v
                else:
                    mix = rng.integers(0, q, N, dtype=np.uint64)
                    mask = rng.random(N) < 0.5
                    mix[mask] = rng.choice(pool, size=int(mask.sum()))
                    out[i, p, l] = mix
    return out


def main() -> None:
    batches = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
    seed_base = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
    threads = os.cpu_count() or 1
    net_pub = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "fhe_precompiles_b200/data/network.pub"), "rb").read()
    _, rk_host = fdev.parse_public_key(net_pub)
    rk = rk_host.cuda()
    rk_np = rk_host.numpy().view(np.uint64)
    mismatched = total = 0
    t0 = time.time()
    for b in range(batches):
        rng = np.random.default_rng(seed_base + b)
        if b % 3 == 2:
            a_np, b_np = adversarial(rng, n), adversarial(rng, n)
        else:
            a_np = np.stack([[rng.integers(0, Q[l], (n, N), dtype=np.uint64) for l in range(2)] for _ in range(2)]).transpose(2, 0, 1, 3).copy()
            b_np = np.stack([[rng.integers(0, Q[l], (n, N), dtype=np.uint64) for l in range(2)] for _ in range(2)]).transpose(2, 0, 1, 3).copy()
        ta = torch.from_numpy(a_np.view(np.int64)).cuda()
        tb = torch.from_numpy(b_np.view(np.int64)).cuda()
        got = fdev.mul_relin(ta, tb, rk).cpu().numpy().view(np.uint64)
        want, _ = bfv.batch_mul_relin(a_np, b_np, rk_np, threads)
        bad = int((got != want).reshape(n, -1).any(axis=1).sum())
        mismatched += bad
        total += n
        print(f"batch {b} ({'adversarial' if b % 3 == 2 else 'uniform'}): {n - bad}/{n} identical", flush=True)
    print(json.dumps({"ops": total, "mismatched_ops": mismatched, "coefficients_compared": total * 4 * N,
                      "batches": batches, "adversarial_batches": batches // 3, "seed_base": seed_base, "seconds": round(time.time() - t0, 1)}))


if __name__ == "__main__":
    main()
