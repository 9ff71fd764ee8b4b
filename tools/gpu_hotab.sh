#!/bin/bash
# hot-line cache decision A/B: count-mode scout (product) against the weight-share rule (MMCB_HOT_BYWEIGHT=1) and no cache (MMCB_HOTCACHE=-1)
# usage: tools/gpu_hotab.sh TAG "workload:method ..."
O=gpurun_out; mkdir -p $O; TAG=${1:-hotab}; WLS=${2:-"sphshells:grid sphshells:elem cube60:elem cube60:grid cube60:havel headatlas:elem skinvessel:grid"}
for wl in $WLS; do
  for mode in count weight off; do
    case $mode in count) E="";; weight) E="MMCB_HOT_BYWEIGHT=1";; off) E="MMCB_HOTCACHE=-1";; esac
    env $E MMCB_TRACE=1 python bench.py --workload ${wl%%:*} --method ${wl##*:} --no-cpu-baseline --no-e2e --no-ref-cuda --steps 3 --warmup 2 2>$O/hotab_err.txt | python -c "
import sys,json
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=j['roofline']
print(json.dumps(dict(mode='$mode', workload='$wl', photons_per_ms=round(j['value']), kernel_ms=round(r['kernel_ms'],2), absorbed=round(j['config']['absorbed_fraction'],5))))"
    grep -h "hot-line\|count-mode" $O/hotab_err.txt | sort -u | head -3
  done
done 2>&1 | tee $O/hotab_${TAG}.log
