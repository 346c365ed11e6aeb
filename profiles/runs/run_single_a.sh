set -x
cd /root/repo
python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r2_pytest5.txt; cat gpurun_out/r2_pytest5.txt
python bench.py > gpurun_out/r2_bench_1gpu_d.json 2> gpurun_out/r2_bench_1gpu_d.err
python3 -c "
import json
d=json.loads(open('gpurun_out/r2_bench_1gpu_d.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['stage_ms_per_step'])"
tail -3 gpurun_out/r2_bench_1gpu_d.err
