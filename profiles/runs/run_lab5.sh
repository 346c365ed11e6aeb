set -x
python profiles/kernel_lab.py --libs default,pre6,pre7,bwd5 --steps 30 > gpurun_out/r2_lab5.jsonl 2> gpurun_out/r2_lab5.err
cat gpurun_out/r2_lab5.jsonl
rm -f gpurun_out/parity_stats.jsonl
python -m pytest tests -m gpu -q -rA 2>&1 | tail -150 > gpurun_out/r2_pytest3.txt
tail -12 gpurun_out/r2_pytest3.txt
