set -x
cd /root/repo
for i in 1 2 3; do
  python bench.py --no-cpu > gpurun_out/r2_bench_repeat_$i.json 2> gpurun_out/r2_bench_repeat_$i.err
  python3 -c "
import json
d=json.loads(open('gpurun_out/r2_bench_repeat_$i.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['host_step_ms'])"
done
SURFEL_BENCH_NOCLOCKS=1 python bench.py --no-cpu --no-e2e > gpurun_out/r2_bench_repeat_noclocks.json 2>/dev/null
python3 -c "
import json
d=json.loads(open('gpurun_out/r2_bench_repeat_noclocks.json').read().strip().splitlines()[-1])
print('noclocks', d['value'], d['ms_per_step'], d['host_step_ms'])"
