#!/bin/bash
# round-end measurement campaign on one B200: one bench line per workload (value, e2e, roofline, reference CUDA on the same workload)
O=gpurun_out; mkdir -p $O; TAG=${1:-r2}
: > $O/bench_${TAG}_workloads.jsonl
for wl in ${WLS:-sphshells:grid sphshells:elem cube60:elem cube60:grid cube60:havel cube60:plucker skinvessel:grid headatlas:elem headlike:elem}; do
  extra=""
  case $wl in cube60:havel|cube60:plucker) extra="--no-ref-cuda";; esac     # the reference's GPU path has no Havel / Plucker tracer
  timeout 900 python bench.py --workload ${wl%%:*} --method ${wl##*:} --steps 3 --warmup 3 --no-cpu-baseline $extra 2>>$O/bench_${TAG}.err | tail -1 >> $O/bench_${TAG}_workloads.jsonl
  tail -1 $O/bench_${TAG}_workloads.jsonl | python -c "
import sys,json
j=json.loads(sys.stdin.read()); r=j['roofline']; rc=j.get('ref_cuda') or {}
print('$wl', 'value %.0f ph/ms' % j['value'], 'kernel %.2f ms' % r['kernel_ms'], 'e2e %.0f (%.1f ms)' % (j['e2e']['value'], j['e2e']['ms']), 'bound', r['bound'], 'frac %.3f' % r['frac'], 'ref_cuda %s ms' % rc.get('kernel_ms'), 'vs_ref_cuda', j.get('vs_ref_cuda'))"
done
