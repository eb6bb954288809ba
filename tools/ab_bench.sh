#!/bin/bash
# A/B of two builds of libmcgvc.so on ONE box (boxes differ by ~2 % under the power cap):
#   tools/ab_bench.sh <other.so> [rounds]    -> ms/step of (other, in-tree) alternating
other=$1; rounds=${2:-3}
for i in $(seq $rounds); do
  a=$(MCGVC_LIBRARY=$other python bench.py --steps 15 --warmup 4 --no-cpu --fast-steps 0 --profile-steps 0 2>/dev/null | tail -1 | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("%.2f %s" % (d["ms_per_step"], d["clocks"]["sm_mhz"]))')
  b=$(python bench.py --steps 15 --warmup 4 --no-cpu --fast-steps 0 --profile-steps 0 2>/dev/null | tail -1 | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("%.2f %s" % (d["ms_per_step"], d["clocks"]["sm_mhz"]))')
  echo "round $i: other $a | in-tree $b"
done
