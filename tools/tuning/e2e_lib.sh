#!/bin/bash
# usage: e2e_lib.sh <lib.so> [env...] : swap the library in, run the end-to-end probe with the debug timeline
lib=$1; shift
cp integrated_path_planning_b200/libfot.so /tmp/libfot_saved.so
cp $lib integrated_path_planning_b200/libfot.so
echo "== $lib $*"
env "$@" FOT_DEBUG_TIMING=1 python tools/tuning/e2e_probe.py 4096 8 2>&1 | grep -v "pointer attr" | tail -5
cp /tmp/libfot_saved.so integrated_path_planning_b200/libfot.so
