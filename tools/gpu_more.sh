#!/bin/bash
# other BASELINE configs (C3 skinvessel, C4 head-like stand-in) and the tail share (1e7 vs 1e8 photons); outputs in gpurun_out/
O=gpurun_out; mkdir -p $O; TAG=${1:-m}
for wl in "skinvessel grid 1e7" "headlike elem 1e7" "cube60 elem 1e8" "cube60 elem 1e6" "sphshells grid 1e8" "sphshells grid 1e6"; do
  set -- $wl
  timeout 600 python bench.py --workload $1 --method $2 --photons $3 --no-cpu-baseline --steps 2 --warmup 3 2>> $O/more_${TAG}.err | python -c "
import sys,json
j=json.loads(sys.stdin.read()); r=j['roofline']
print(json.dumps(dict(workload='$1:$2:$3', photons_per_ms=round(j['value']), kernel_ms=round(r['kernel_ms'],2), step_ms=round(j['ms_per_step'],2), e2e=j.get('e2e'), steps_per_photon=round(j['config']['raytet_steps_per_photon'],1), gsteps_s=round(r['gsteps_per_s'],1), absorbed=round(j['config']['absorbed_fraction'],5))))"
done 2>&1 | tee $O/more_${TAG}.log
MMCB_TRACE=1 python bench.py --workload headlike --method elem --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2> $O/trace_headlike_${TAG}.log
