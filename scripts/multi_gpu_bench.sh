N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_mg${N}_c2.json 2> gpurun_out/r2_mg${N}_c2.err
grep -h '^{' gpurun_out/r2_mg${N}_c2.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['n_gpus'], round(d['value']), round(d['e2e']['value']), d['ms_per_step'], d['e2e']['ms_per_step'])"
