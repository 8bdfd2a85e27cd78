#!/bin/bash
# how fast is the one-kernel-per-phase path when the whole batch is L2-resident?  (step time x (64 / batch) = what batch chunking could reach)
for b in 8 16 32 64; do
  for mode in "--force-split" ""; do
    python bench.py --batch $b $mode --steps 200 --warmup 10 --no-cpu --no-variant --no-e2e --no-workloads 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('batch $b ${mode:-cluster}', 'ms', d['ms_per_step'], 'x(64/b)', round(d['ms_per_step']*64/$b,4), 'frac', d['roofline']['step_frac'], 'launches', d['launches_per_step'])"
  done
done
