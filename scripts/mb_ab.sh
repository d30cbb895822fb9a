#!/bin/bash
# A/B of the forward-pass slice size on the headline config (alternating runs on one box)
for mb in 256 128 256 128; do
  python bench.py --steps 4 --warmup 3 --other-modes "" --no-cpu-baseline --no-hbm-kernels --check-samples 1 --microbatch $mb 2>/dev/null > /tmp/mb.json
  python -c "import json; d=json.load(open('/tmp/mb.json')); print('microbatch $mb', round(d['value'],1), round(d['e2e']['value'],1), d['clocks']['sm_mhz'], d['config']['workspace_bytes'])"
done
