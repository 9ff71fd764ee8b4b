#!/bin/bash
# ncu full captures of the photon-kernel variants that are not the headline: Havel on cube60 (the C1 tracer) and the detector kernel on the
# head-like mesh (C4 stand-in), at the benched 1e7 photons; outputs in gpurun_out/
set -u
O=gpurun_out; mkdir -p $O; TAG=${1:-x}
timeout 240 ncu --set full --clock-control none --import-source on -k regex:mmcb_photon -s 1 -c 1 -f -o $O/prof_havel_${TAG} \
    python bench.py --workload cube60 --method havel --steps 1 --warmup 1 --photons 1e7 --no-cpu-baseline --no-e2e > $O/ncu_havel_${TAG}.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mmcb_photon -s 1 -c 1 -f -o $O/prof_headdet_${TAG} \
    python bench.py --workload headlike --steps 1 --warmup 1 --photons 1e7 --no-cpu-baseline --no-e2e > $O/ncu_headdet_${TAG}.log 2>&1
ls -la $O | grep ${TAG}
