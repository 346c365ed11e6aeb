set -x
cd /root/repo
rm -f gpurun_out/parity_stats.jsonl
python -m pytest tests -q -m gpu -rf 2>&1 | tail -15 > gpurun_out/r2_pytest7.txt; tail -6 gpurun_out/r2_pytest7.txt
python bench.py --no-cpu > gpurun_out/r2_bench_1gpu_e.json 2> gpurun_out/r2_bench_1gpu_e.err
python3 -c "
import json
d=json.loads(open('gpurun_out/r2_bench_1gpu_e.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['stage_ms_per_step'])"
tail -3 gpurun_out/r2_bench_1gpu_e.err
