"""Multi-GPU path on CPU: the batch shards across ranks with no data-path collective; world_size-2 gloo run checks
the slices, the max-over-ranks timing and the whole-job rate the benchmark reports."""
import os
import socket
import subprocess
import sys

from helpers import ROOT


def test_shard_range_partitions_exactly():
    from fhe_precompiles_b200.sharding import shard_range

    for n in (0, 1, 7, 4096, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_cost_weighted_sharding_balances_mixed_batches():
    from fhe_precompiles_b200 import _lib
    from fhe_precompiles_b200.sharding import call_cost, shard_by_cost

    import random

    rnd = random.Random(3)
    names = [rnd.choice(_lib.BINARY_OPS) for _ in range(4096)]
    costs = [call_cost(n) for n in names]
    assert call_cost("mul_cipheri64_cipheri64") > 5 * call_cost("mul_cipheri64_i64") > 5 * call_cost("add_cipheri64_cipheri64")
    for world in (2, 4, 8):
        parts = shard_by_cost(costs, world)
        assert sorted(i for p in parts for i in p) == list(range(len(names)))
        loads = [sum(costs[i] for i in p) for p in parts]
        assert max(loads) / (sum(loads) / world) < 1.01


WORKER = r"""
import os, sys, time
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as dist
from fhe_precompiles_b200.sharding import shard_range, max_over_ranks, whole_job_rate
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
lo, hi = shard_range(4097, rank, world)
items = torch.arange(lo, hi, dtype=torch.int64)
# every rank processes only its slice; the only communication is the barrier and the max of the elapsed time
dist.barrier()
elapsed = 0.010 * (rank + 1)
slowest = max_over_ranks(elapsed, dist)
counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
dist.all_gather(counts, torch.tensor([items.numel()]))
if rank == 0:
    total = int(sum(c.item() for c in counts))
    print("RESULT", total, round(slowest, 3), round(whole_job_rate(100, world, slowest), 1), flush=True)
dist.destroy_process_group()
"""


def test_two_rank_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script), ROOT]
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT")][0].split()
    assert int(line[1]) == 4097  # slices cover the batch exactly once
    assert float(line[2]) == 0.02  # max over ranks
    assert float(line[3]) == 10000.0  # 2 ranks x 100 units / 0.02 s
