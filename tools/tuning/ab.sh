#!/bin/bash
# usage: ab.sh "ENV1=.. ENV2=.." "ENV.." ... : resident + e2e bench (--brief) per environment, same box, prints stage ms
for v in "$@"; do
  env $v python bench.py --steps 20 --warmup 3 --brief > /tmp/ab.json 2>/tmp/ab.err
  python -c "
import json; d=json.load(open('/tmp/ab.json')); print('$v', {k: round(x, 3) for k, x in d['roofline']['stage_ms'].items()}, 'frac', round(d['roofline']['frac'], 3), 'step', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['ms_per_step'], 3), d['clocks']['sm_mhz'])" || tail -3 /tmp/ab.err
done
