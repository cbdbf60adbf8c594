#!/bin/bash
# N = 1, 2, 4, 8 of bench.py back to back on ONE 8-GPU box (what the driver's scaling run does), short form:
# per-N ms_per_step, value, per-rank times.  gpurun --gpus 8 -- 'bash tools/scale_same_box.sh'
for N in 1 2 4 8; do
  if [ $N = 1 ]; then CMD="python bench.py"; else CMD="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N"; fi
  $CMD --steps 50 --warmup 5 --reps 9 --no-extras --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json
l=[x for x in sys.stdin if x.startswith('{')][-1]; d=json.loads(l)
print(json.dumps({'n_gpus': d['n_gpus'], 'ms_per_step': round(d['ms_per_step'],5), 'value_T': round(d['value']/1e6,3), 'k2_frac_rank0': round(d['roofline']['frac'],4), 'per_rank_ms': d['repetitions']['per_rank_ms_per_step_median_rep'], 'host_enqueue_ms': d['repetitions'].get('host_enqueue_ms_per_step_median_rep_per_rank')}))
"
done
