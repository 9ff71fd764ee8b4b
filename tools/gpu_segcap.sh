#!/bin/bash
# sweep of MMCB_SEGCAP (voxel edges of path a dual-grid lane deposits per iteration; 0 = CAP kernel variant off) on the grid workloads
O=gpurun_out; mkdir -p $O
for wl in ${WLS:-skinvessel:grid sphshells:grid cube60:grid}; do
  for cap in ${CAPS:-0 2 3 4 6 8 12}; do
    MMCB_SEGCAP=$cap python bench.py --workload ${wl%%:*} --method ${wl##*:} --no-cpu-baseline --no-e2e --no-ref-cuda --steps 3 --warmup 2 2>/dev/null | python -c "
import sys,json
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=j['roofline']
print(json.dumps(dict(segcap=$cap, workload='$wl', photons_per_ms=round(j['value']), kernel_ms=round(r['kernel_ms'],2), absorbed=round(j['config']['absorbed_fraction'],5), steps_per_photon=round(j['config']['raytet_steps_per_photon'],2))))"
  done
done 2>&1 | tee $O/segcap.log
