set -x
python profiles/kernel_lab.py --libs default,bb352 --steps 30 > gpurun_out/r2_lab4.jsonl 2> gpurun_out/r2_lab4.err
cat gpurun_out/r2_lab4.jsonl
rm -f gpurun_out/parity_stats.jsonl
python -m pytest tests -m gpu -q -rA 2>&1 | tail -150 > gpurun_out/r2_pytest2.txt
tail -15 gpurun_out/r2_pytest2.txt
