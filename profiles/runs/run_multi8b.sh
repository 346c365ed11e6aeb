set -x
cd /root/repo
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $R --master-port 29517 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu > gpurun_out/r2_bench_8gpu_b.json 2> gpurun_out/r2_bench_8gpu_b.err
python3 -c "
import json
d=json.loads(open('gpurun_out/r2_bench_8gpu_b.json').read().strip().splitlines()[-1])
t=d.get('tile_band') or {}; [t.pop(k,None) for k in ('how','overlapped','fused')]
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], json.dumps(t))"
tail -3 gpurun_out/r2_bench_8gpu_b.err
