#!/bin/bash
# GPU box: parity tests, then A/B of the lane re-packing kernel against the flattened one (MMCB_NO_REPACK=1)
# usage: tools/gpu_rp.sh TAG [pytest-args]
O=gpurun_out; mkdir -p $O; TAG=${1:-rp}; shift
if [ -z "$NOTEST" ]; then timeout 1500 python -m pytest tests -m gpu -x -q "$@" > $O/pytest_${TAG}.log 2>&1; echo "pytest exit $?" >> $O/pytest_${TAG}.log
tail -5 $O/pytest_${TAG}.log; fi
WLS=${WLS:-"sphshells:grid sphshells:elem cube60:elem cube60:grid headlike:elem"}
for wl in $WLS; do
  for norp in 1 0; do
    if [ $norp = 1 ]; then unset MMCB_REPACK; else export MMCB_REPACK=1; fi
    timeout 600 python bench.py --workload ${wl%%:*} --method ${wl##*:} --no-cpu-baseline --no-e2e --steps 3 --warmup 2 2>$O/err_${TAG}.log | python -c "
import sys,json
try:
    j=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=j['roofline']
    print(json.dumps(dict(norepack=$norp, workload='$wl', photons_per_ms=round(j['value']), kernel_ms=round(r['kernel_ms'],2), absorbed=round(j['config']['absorbed_fraction'],5), steps_per_photon=round(j['config']['raytet_steps_per_photon'],2))))
except Exception as e:
    print('FAILED $wl norepack=$norp', e); print(open('$O/err_${TAG}.log').read()[-1500:])"
  done
done 2>&1 | tee $O/ab_${TAG}.log
