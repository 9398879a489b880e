#!/bin/bash
# usage: bench_lib.sh <lib.so> [env assignments...] : swap the library in, run the resident bench, print sweep ms
lib=$1; shift
cp integrated_path_planning_b200/libfot.so /tmp/libfot_saved.so
cp $lib integrated_path_planning_b200/libfot.so
env "$@" python bench.py --steps 10 --warmup 3 --no-cpu > /tmp/b.json 2>/tmp/b.err
python -c "
import json; d=json.load(open('/tmp/b.json')); print('$lib $*', d['roofline']['stage_ms'], d['roofline']['frac'])" || tail -3 /tmp/b.err
cp /tmp/libfot_saved.so integrated_path_planning_b200/libfot.so
