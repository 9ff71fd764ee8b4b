#!/bin/bash
# BASELINE config C4 strong scaling: head atlas mesh, detectors + partial paths, PHOTONS (default 1e9) photons over N GPUs; usage: tools/gpu_scale_c4.sh N [photons]
O=gpurun_out; mkdir -p $O; N=$1; PH=${2:-1e9}
if [ "$N" = 1 ]; then
  python bench.py --gpus 1 --workload headatlas --scaling strong --photons $PH --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-ref-cuda > $O/scale_c4_n$N.json 2> $O/scale_c4_n$N.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload headatlas --scaling strong --photons $PH --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-ref-cuda > $O/scale_c4_n$N.json 2> $O/scale_c4_n$N.err
fi
tail -c 900 $O/scale_c4_n$N.json; tail -3 $O/scale_c4_n$N.err
