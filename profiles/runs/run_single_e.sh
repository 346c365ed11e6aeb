set -x
cd /root/repo
for i in 1 2 3; do
  python bench.py --no-cpu --no-e2e > gpurun_out/r2_bench_nvml_$i.json 2> gpurun_out/r2_bench_nvml_$i.err
done
for i in 1 2; do
  python bench.py --no-cpu --no-e2e --steps 20 --warmup 5 > gpurun_out/r2_bench_nvml_s20_$i.json 2> gpurun_out/r2_bench_nvml_s20_$i.err
done
SURFEL_BENCH_CLOCKS=smi python bench.py --no-cpu --no-e2e > gpurun_out/r2_bench_smi200.json 2>/dev/null
python3 -c "
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench_nvml*.json'))+['gpurun_out/r2_bench_smi200.json']:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(d['value'],1), round(d['ms_per_step'],4), {k:round(v,3) for k,v in d['host_step_ms'].items() if k!='what'}, d['clocks'])"
tail -3 gpurun_out/r2_bench_nvml_1.err
