#!/bin/bash
# Runs on the GPU box (under gpurun): bench lines for the three kernels' workloads, the ncu launch list of the default
# bench command and one `--set full` capture per photon-kernel variant.  Outputs land in gpurun_out/.
set -u
O=gpurun_out
mkdir -p $O
TAG=${1:-r1}
python bench.py > $O/bench_${TAG}.json 2> $O/bench_${TAG}.err
for wl in "sphshells elem" "cube60 elem" "cube60 grid"; do
  set -- $wl
  python bench.py --workload $1 --method $2 --no-cpu-baseline --steps 3 --warmup 3 >> $O/bench_${TAG}_more.json 2>> $O/bench_${TAG}.err
done
MMCB_TRACE=1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2> $O/trace_${TAG}.log
# launch list of the default bench command (cold-cache, serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $O/ncu_launch_${TAG}.log 2>&1
# full captures at the benched photon count (about 40 replays of a 30-170 ms kernel)
ncu --set full --clock-control none --import-source on -k regex:mmcb_photon -s 1 -c 1 -f -o $O/prof_grid_${TAG} \
    python bench.py --steps 1 --warmup 1 --photons 1e7 --no-cpu-baseline --no-e2e > $O/ncu_grid_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mmcb_photon -s 1 -c 1 -f -o $O/prof_elem_${TAG} \
    python bench.py --workload cube60 --method elem --steps 1 --warmup 1 --photons 1e7 --no-cpu-baseline --no-e2e > $O/ncu_elem_${TAG}.log 2>&1
ls -la $O
