#!/bin/bash
# ncu --set full captures of the photon kernel for the bench workloads (1e7 photons); reports land in gpurun_out/prof_<tag>_<workload>.ncu-rep
O=gpurun_out; mkdir -p $O; TAG=${1:-r2}
WLS=${WLS:-"sphshells:grid cube60:elem skinvessel:grid headatlas:elem"}
for wl in $WLS; do
  name=${wl%%:*}; m=${wl##*:}
  ncu --set full --clock-control none --import-source on -k regex:mmcb_photon -s 2 -c 1 -f -o $O/prof_${TAG}_${name}_${m} \
      python bench.py --workload $name --method $m --steps 1 --warmup 1 --photons ${PHOTONS:-1e7} --no-cpu-baseline --no-e2e --no-ref-cuda > $O/ncu_${TAG}_${name}.log 2>&1
  ls -la $O/prof_${TAG}_${name}_${m}.ncu-rep
done
