set -x
python profiles/determinism_probe.py --iters 400 > gpurun_out/r2_determinism.jsonl 2> gpurun_out/r2_determinism.err
SURFEL_LIB=$PWD/2d-gaussian-splatting_b200/lib/variants/pre6.so python profiles/determinism_probe.py --iters 400 >> gpurun_out/r2_determinism.jsonl 2>> gpurun_out/r2_determinism.err
python profiles/determinism_probe.py --iters 300 --workload config2 >> gpurun_out/r2_determinism.jsonl 2>> gpurun_out/r2_determinism.err
cat gpurun_out/r2_determinism.jsonl; tail -3 gpurun_out/r2_determinism.err
