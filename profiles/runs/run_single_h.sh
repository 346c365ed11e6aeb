set -x
cd /root/repo
python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2_s20_gcpaused.json 2>/dev/null
SURFEL_BENCH_E2E_KEEP_GC=1 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2_s20_e2e_keepgc.json 2>/dev/null
SURFEL_BENCH_KEEP_GC=1 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2_s20_keepgc.json 2>/dev/null
python3 -c "
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_s20_*.json')):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(d['value'],1), round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],3), d['host_step_ms']['max'])"
