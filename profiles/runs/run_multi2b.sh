set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e > gpurun_out/r2_bench_2gpu_b.json 2> gpurun_out/r2_bench_2gpu_b.err
python3 -c "
import json
d=json.loads(open('gpurun_out/r2_bench_2gpu_b.json').read().strip().splitlines()[-1])
print(json.dumps(d.get('tile_band'),indent=1))"
tail -3 gpurun_out/r2_bench_2gpu_b.err
