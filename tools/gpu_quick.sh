#!/bin/bash
# quick GPU check: parity tests, then one short bench line per workload/kernel variant (product library)
O=gpurun_out; mkdir -p $O; TAG=${1:-q}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for wl in "sphshells grid" "sphshells elem" "cube60 elem" "cube60 grid"; do
  set -- $wl
  python bench.py --workload $1 --method $2 --no-cpu-baseline --no-e2e --steps 3 --warmup 2 | python -c "
import sys,json
j=json.loads(sys.stdin.read()); r=j['roofline']
print(json.dumps(dict(workload='$1:$2', photons_per_ms=round(j['value']), kernel_ms=round(r['kernel_ms'],2), step_ms=round(j['ms_per_step'],2), gsteps_s=round(j['config']['raytet_steps_per_photon']*1e7/r['kernel_ms']/1e6,1), absorbed=round(j['config']['absorbed_fraction'],5))))"
done 2>&1 | tee $O/quick_${TAG}.log
