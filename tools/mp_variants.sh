#!/bin/bash
# A/B of the line-by-line kernel's geometry behind the far-field pass (bench.py configs[3] shard, 3 steps)
for v in 10 11 0 3; do
  AB200_SUM_MP_VARIANT=$v python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('variant $v', 'ms/step', round(d['ms_per_step'],1), 'evals/s %.3e'%d['value'], 'kernel_ms', {k:round(x,2) for k,x in d['kernel_ms'].items()}, 'checksum', d['checksum_I'])"
done
