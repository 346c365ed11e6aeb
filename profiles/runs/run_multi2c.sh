set -x
cd /root/repo
python -m pytest tests/test_parity_gpu.py -q -k "out_buffers" 2>&1 | tail -3
python -m pytest tests/test_multigpu_gpu.py -q 2>&1 | tail -3
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$R --master-port 29533 profiles/band_probe.py > gpurun_out/r2_band_probe_fixed.json 2> gpurun_out/r2_band_probe_fixed.err
tail -c 1200 gpurun_out/r2_band_probe_fixed.json
$R --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e > gpurun_out/r2_bench_2gpu_c.json 2> gpurun_out/r2_bench_2gpu_c.err
python3 -c "
import json
d=json.loads(open('gpurun_out/r2_bench_2gpu_c.json').read().strip().splitlines()[-1])
t=d.get('tile_band'); t.pop('how',None); print(json.dumps(t))"
tail -3 gpurun_out/r2_bench_2gpu_c.err
