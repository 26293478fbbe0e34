#!/bin/bash
# byte-surface batch throughput against FHE_B200_TILE_OPS (run on the GPU box from the repo root)
for t in 4 8 32 64; do
  FHE_B200_TILE_OPS=$t python bench.py --no-cpu-baseline 2>/dev/null | T=$t python -c '
import json,sys,os
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); l=d["latency"]["byte_surface_batch"]
print("tile", os.environ["T"], round(l["calls_per_s"]), round(l["chained_calls_per_s"]))'
done
