set -x
cd /root/repo
rm -f gpurun_out/parity_stats.jsonl
python -m pytest tests -q -m gpu -rf 2>&1 | tail -40 > gpurun_out/r2_pytest6.txt; tail -12 gpurun_out/r2_pytest6.txt
