set -x
python profiles/kernel_lab.py --libs default,default,pre6,pre6,pre7 --steps 10 > gpurun_out/r2_lab6.jsonl 2> gpurun_out/r2_lab6.err
cat gpurun_out/r2_lab6.jsonl
for tool in memcheck racecheck initcheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python profiles/sanitizer_workload.py > gpurun_out/r2_sanitizer_$tool.txt 2>&1
  tail -4 gpurun_out/r2_sanitizer_$tool.txt
done
SURFEL_LIB=$PWD/2d-gaussian-splatting_b200/lib/variants/pre6.so timeout 600 compute-sanitizer --tool initcheck --print-limit 20 python profiles/sanitizer_workload.py > gpurun_out/r2_sanitizer_initcheck_pre6.txt 2>&1
tail -4 gpurun_out/r2_sanitizer_initcheck_pre6.txt
rm -f gpurun_out/parity_stats.jsonl
python -m pytest tests -m gpu -q -rA 2>&1 | tail -150 > gpurun_out/r2_pytest4.txt
tail -12 gpurun_out/r2_pytest4.txt
