#!/bin/bash
# hot-line cache on/off per workload + the scout's hottest-line share (decides MMCB_HOT_MINSHARE); output gpurun_out/hot_<tag>.log
O=gpurun_out; mkdir -p $O; TAG=${1:-h}
for wl in "sphshells grid" "sphshells elem" "cube60 elem" "cube60 grid" "skinvessel grid" "headlike elem"; do
  set -- $wl
  for hot in -1 1; do
    MMCB_TRACE=1 MMCB_HOTCACHE=$hot python bench.py --workload $1 --method $2 --no-cpu-baseline --no-e2e --steps 3 --warmup 2 2> $O/hot_trace.tmp | python -c "
import sys,json
j=json.loads(sys.stdin.read()); r=j['roofline']
print(json.dumps(dict(workload='$1:$2', hot=$hot, photons_per_ms=round(j['value']), kernel_ms=round(r['kernel_ms'],2))))"
    grep -h "hottest line" $O/hot_trace.tmp | tail -1
  done
done 2>&1 | tee $O/hot_${TAG}.log
