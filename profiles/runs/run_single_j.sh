set -x
cd /root/repo
python -m pytest tests/test_parity_gpu.py -q -k "replicated_output" 2>&1 | tail -8
