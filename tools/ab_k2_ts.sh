#!/bin/bash
# same-box A/B of the band-power kernel: shared-memory A operand (MS_K2_TS=0) against A through tensor memory (1),
# optionally crossed with another knob:  tools/ab_k2_ts.sh [VAR "v1 v2 ..."]
VAR=${1:-MS_UNUSED}; VALS=${2:-x}
for i in 1 2; do
for val in $VALS; do
for v in 0 1; do
env $VAR=$val MS_K2_TS=$v python bench.py --steps 50 --warmup 5 --reps 7 --no-extras --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json
l=[x for x in sys.stdin if x.startswith('{')][-1]; d=json.loads(l)
print('$VAR=$val TS=$v', 'ms_per_step', round(d['ms_per_step'],5), 'k2_ms', round(d['roofline']['kernel_ms'],5), 'frac', round(d['roofline']['frac'],4), d['hourly_counts']['anzahl_total'])
"
done; done; done
