#!/bin/bash
# A/B a tuning environment variable of the library on ONE box: tools/ab_env.sh VAR "v1 v2 ..." [bench args]
var=$1; vals=$2; shift 2
for v in $vals; do
  env $var=$v python bench.py --steps 300 --warmup 20 --no-e2e "$@" 2>/dev/null | tail -1 > /tmp/ab.json
  python - "$var" "$v" <<'PY'
import json, sys
d = json.load(open("/tmp/ab.json"))
dl = d.get("dense_layout") or {}
print(sys.argv[1], sys.argv[2], "step_ms", round(d["ms_per_step"], 5), "frac", round(d["roofline"]["frac"], 4),
      "dense_kernel_ms", dl.get("kernel_ms"), "dense_frac", dl.get("roofline_frac"))
PY
done
