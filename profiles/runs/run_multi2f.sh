set -x
cd /root/repo
timeout 300 python -m pytest tests/test_multigpu_gpu.py tests/test_parity_gpu.py -q -s -k "two_gpus or deferred or out_buffers" 2>&1 | tail -8
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $R --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2_bench_2gpu_h.json 2> gpurun_out/r2_bench_2gpu_h.err
python3 -c "
import json
d=json.loads(open('gpurun_out/r2_bench_2gpu_h.json').read().strip().splitlines()[-1])
t=d.get('tile_band'); [t.pop(k,None) for k in ('how','overlapped','fused')]; print(json.dumps(t))"
tail -5 gpurun_out/r2_bench_2gpu_h.err
