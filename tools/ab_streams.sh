#!/bin/bash
# usage: tools/ab_streams.sh <streams> [<streams> ...]  -- ms/step of the pipelined device-resident step per stream count
for st in "$@"; do
  python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-single --step-streams "$st" 2>/dev/null | grep '^{' > /tmp/ab_streams.json
  python - "$st" <<'PY'
import json, sys
d = json.load(open("/tmp/ab_streams.json"))
print("streams", sys.argv[1], "ms/step", round(d["ms_per_step"], 2), "value", round(d["value"]), "sm_mhz", d["clocks"]["sm_mhz"], flush=True)
PY
done
