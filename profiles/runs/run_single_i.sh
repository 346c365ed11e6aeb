set -x
cd /root/repo
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python -m pytest tests/test_parity_gpu.py tests/test_multigpu_gpu.py -q -k "out_buffers or tile_band or deferred or public_api or two_gpus" 2>&1 | tail -3
