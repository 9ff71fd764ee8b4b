#!/bin/bash
# ncu full captures of the photon kernel (grid + elem workloads) for source-line analysis; outputs in gpurun_out/
set -u
O=gpurun_out; mkdir -p $O; TAG=${1:-x}
ncu --set full --clock-control none --import-source on -k regex:mmcb_photon -s 1 -c 1 -f -o $O/prof_grid_${TAG} \
    python bench.py --steps 1 --warmup 1 --photons 1e6 --no-cpu-baseline --no-e2e > $O/ncu_grid_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mmcb_photon -s 1 -c 1 -f -o $O/prof_elem_${TAG} \
    python bench.py --workload cube60 --method elem --steps 1 --warmup 1 --photons 1e6 --no-cpu-baseline --no-e2e > $O/ncu_elem_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mmcb_photon -s 1 -c 1 -f -o $O/prof_selem_${TAG} \
    python bench.py --workload sphshells --method elem --steps 1 --warmup 1 --photons 1e6 --no-cpu-baseline --no-e2e > $O/ncu_selem_${TAG}.log 2>&1
ls -la $O
