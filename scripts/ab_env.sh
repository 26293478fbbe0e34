#!/bin/bash
# usage: scripts/ab_env.sh VAR v1 v2 ...   -> device-resident ops/s and per-kernel us/op of bench.py under VAR=v for each value
VAR=$1; shift
for v in "$@"; do
  env $VAR=$v python bench.py --no-cpu-baseline --no-e2e --no-configs 2>/dev/null | VAL="$VAR=$v" python -c '
import sys, json, os
d = json.loads(sys.stdin.read())
print(os.environ["VAL"], round(d["value"]), {k: round(x["us_per_op"], 4) for k, x in d["roofline"]["per_kernel"].items()})'
done
