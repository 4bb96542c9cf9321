#!/bin/bash
# Where do the far-field sums start to pay?  configs[1] (5 species x --lines-per-species lines, 1e5 frequencies, 100 levels) with
# the line-by-line kernel (AB200_FARFIELD=0) against the far-field sums forced on (AB200_FARFIELD=2), and a configs[4] path.
for lps in 2000 5000 10000 20000; do
  for ff in 0 2; do
    AB200_FARFIELD=$ff python bench.py --workload c2 --lines-per-species $lps --steps 3 --warmup 2 --no-cpu-baseline --no-extra 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('lines/species $lps', 'AB200_FARFIELD=$ff', 'ms/step', round(d['ms_per_step'],2), 'checksum', d['checksum_I'])"
  done
done
