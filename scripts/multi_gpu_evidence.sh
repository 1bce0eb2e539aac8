#!/bin/bash
# Run with gpurun --gpus N: bench.py under torchrun on N GPUs (C2 weak scaling and C3 strong scaling through the
# CLI's partitioner) + the pinned-copy ceiling of the box.  Outputs -> gpurun_out/r2_mg<N>_*.json
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_mg${N}_topo.log 2>&1
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
$RUN bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_mg${N}_c2.json 2> gpurun_out/r2_mg${N}_c2.err
$RUN bench.py --gpus $N --steps 3 --warmup 3 --workload C3 > gpurun_out/r2_mg${N}_c3.json 2> gpurun_out/r2_mg${N}_c3.err
$RUN scripts/host_link_probe.py > gpurun_out/r2_mg${N}_link.json 2> gpurun_out/r2_mg${N}_link.err
grep -h '^{' gpurun_out/r2_mg${N}_c2.json gpurun_out/r2_mg${N}_c3.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['config']['workload'][:3], d['n_gpus'], round(d['value']), round(d['e2e']['value']), d['ms_per_step'], d['e2e']['ms_per_step'])"
tail -1 gpurun_out/r2_mg${N}_link.json
