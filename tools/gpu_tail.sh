#!/bin/bash
# photons/ms at small photon counts (tail share of the run): usage tools/gpu_tail.sh "variant ..." "workload:method ..." "photons ..."
O=gpurun_out; mkdir -p $O; VARS=${1:-product}; WLS=${2:-"cube60:elem sphshells:grid"}; NS=${3:-"1e6 1e7"}
for v in $VARS; do
  if [ "$v" = product ]; then LIB=mmc_b200/libmmc_b200.so; else LIB=build/variants/libmmc_b200_$v.so; fi
  for wl in $WLS; do for n in $NS; do
    MMCB_LIB=$PWD/$LIB python bench.py --workload ${wl%%:*} --method ${wl##*:} --photons $n --no-cpu-baseline --no-e2e --no-ref-cuda --steps 5 --warmup 3 2>/dev/null | python -c "
import sys,json
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=j['roofline']
print(json.dumps(dict(variant='$v', workload='$wl', photons=float('$n'), photons_per_ms=round(j['value']), kernel_ms=round(r['kernel_ms'],3))))"
  done; done
done 2>&1 | tee -a $O/tail.log
