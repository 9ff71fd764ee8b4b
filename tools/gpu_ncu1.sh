#!/bin/bash
# one ncu full capture of the main photon kernel launch of a short bench run; usage: tools/gpu_ncu1.sh TAG workload method [photons]
O=gpurun_out; mkdir -p $O; TAG=$1; WL=$2; M=$3; N=${4:-1e6}
ncu --set full --clock-control none --import-source on -k regex:mmcb_photon -s 2 -c 1 -f -o $O/prof_${TAG} \
    python bench.py --workload $WL --method $M --steps 1 --warmup 1 --photons $N --no-cpu-baseline --no-e2e > $O/ncu_${TAG}.log 2>&1
tail -3 $O/ncu_${TAG}.log
ls -la $O/prof_${TAG}.ncu-rep
