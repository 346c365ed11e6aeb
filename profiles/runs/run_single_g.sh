set -x
cd /root/repo
for i in 1 2 3; do
  python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2_bench_s20_$i.json 2> gpurun_out/r2_bench_s20_$i.err
done
SURFEL_BENCH_NOCLOCKS=1 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2_bench_s20_noclocks.json 2>/dev/null
python3 -c "
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench_s20_*.json')):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],3), d['host_step_ms']['max'])"
