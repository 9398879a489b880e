for v in "A=1" "FOT_PAIR_CPQ=1" "FOT_PAIR_CPQ=2" "FOT_PAIR_CPQ=3" "FOT_PAIR_CPQ=4"; do
  env $v python bench.py --steps 30 --warmup 3 --brief --queries 512 > /tmp/ab.json 2>/tmp/ab.err
  python -c "
import json; d=json.load(open('/tmp/ab.json')); print('$v', {k: round(x, 4) for k, x in d['roofline']['stage_ms'].items()}, 'step', round(d['ms_per_step'], 4))" || tail -3 /tmp/ab.err
done
