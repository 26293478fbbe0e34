#!/bin/bash
# usage: scripts/ab_sustained.sh "VAR=v [VAR2=w]" ...  -> timed-region and sustained (6 s, clocks under load) ops/s of bench.py per setting
for setting in "$@"; do
  env $setting python bench.py --no-cpu-baseline --no-e2e --no-configs --sustain-seconds 6 2>/dev/null | SETTING="$setting" python -c '
import sys, json, os
d = json.loads(sys.stdin.read())
s = d.get("sustained") or {}
print(os.environ["SETTING"], "| timed", round(d["value"]), "| sustained", round(s.get("ops_per_s_per_gpu", 0)), s.get("clocks"))'
done
