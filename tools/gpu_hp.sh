#!/bin/bash
# Havel / Plucker check on the GPU box: parity tests that touch the two tracers, then short bench lines (product or variant libraries)
# usage: tools/gpu_hp.sh TAG "variant ..." [notest]
O=gpurun_out; mkdir -p $O; TAG=${1:-hp}; VARS=${2:-product}
if [ "$3" != notest ]; then
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_c1_1e8.py tests/test_prep_gpu.py tests/test_cli_dropin.py -m gpu -x -q -k "havel or plucker or Havel or Plucker or hp or nodal" 2>&1 | tail -8 | tee $O/hp_${TAG}_tests.log
fi
for v in $VARS; do
  if [ "$v" = product ]; then LIB=mmc_b200/libmmc_b200.so; else LIB=build/variants/libmmc_b200_$v.so; fi
  for wl in ${WLS:-"cube60:havel:0 cube60:plucker:0 sphshells:havel:0 sphshells:plucker:0 cube60:havel:1 cube60:plucker:1 sphshells:havel:1"}; do
    set -- ${wl//:/ }
    MMCB_LIB=$PWD/$LIB python bench.py --workload $1 --method $2 --basisorder $3 --no-cpu-baseline --no-e2e --no-ref-cuda --steps 3 --warmup 2 2>/dev/null | python -c "
import sys,json
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=j['roofline']
print(json.dumps(dict(variant='$v', workload='$1:$2:b$3', photons_per_ms=round(j['value']), kernel_ms=round(r['kernel_ms'],2), absorbed=round(j['config']['absorbed_fraction'],5), steps=round(j['config']['raytet_steps_per_photon'],2))))"
  done
done 2>&1 | tee $O/hp_${TAG}.log
