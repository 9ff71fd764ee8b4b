#!/bin/bash
# A/B of library variants under build/variants (GPU box): one short bench line per variant and workload
# usage: tools/gpu_ab.sh TAG "variant ..." "workload:method ..."
O=gpurun_out; mkdir -p $O; TAG=${1:-ab}; VARS=${2:-"base"}; WLS=${3:-"sphshells:grid"}
for rep in $(seq 1 ${REPS:-2}); do
for v in $VARS; do
  if [ "$v" = product ]; then LIB=mmc_b200/libmmc_b200.so; else LIB=build/variants/libmmc_b200_$v.so; fi
  for wl in $WLS; do
    MMCB_LIB=$PWD/$LIB python bench.py --workload ${wl%%:*} --method ${wl##*:} --no-cpu-baseline --no-e2e --steps 3 --warmup 2 | python -c "
import sys,json
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=j['roofline']
print(json.dumps(dict(variant='$v', rep=$rep, workload='$wl', photons_per_ms=round(j['value']), kernel_ms=round(r['kernel_ms'],2), absorbed=round(j['config']['absorbed_fraction'],5))))"
  done
done
done 2>&1 | tee $O/ab_${TAG}.log
