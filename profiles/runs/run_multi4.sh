set -x
cd /root/repo
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
$R --master-port 29519 profiles/pcie_concurrent.py > gpurun_out/r2_pcie_4.json 2> gpurun_out/r2_pcie_4.err
cat gpurun_out/r2_pcie_4.json
$R --master-port 29517 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2_bench_4gpu.json 2> gpurun_out/r2_bench_4gpu.err
$R --master-port 29521 bench.py --gpus 4 --steps 20 --warmup 5 --no-tile-band --no-cpu --pin-policy interleave > gpurun_out/r2_bench_4gpu_interleave.json 2> gpurun_out/r2_bench_4gpu_interleave.err
python3 -c "
import json
for n in ('r2_bench_4gpu','r2_bench_4gpu_interleave'):
    d=json.loads(open('gpurun_out/'+n+'.json').read().strip().splitlines()[-1])
    t=d.get('tile_band') or {}; t.pop('how',None); t.pop('overlapped',None)
    print(n, d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], json.dumps(t))"
tail -3 gpurun_out/r2_bench_4gpu.err
