set -x
python -m pytest tests/test_multigpu_gpu.py -q -rA 2>&1 | tail -15 > gpurun_out/r2_pytest_2gpu.txt
tail -5 gpurun_out/r2_pytest_2gpu.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err
tail -c 3000 gpurun_out/r2_bench_2gpu.json; tail -5 gpurun_out/r2_bench_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 profiles/pcie_concurrent.py > gpurun_out/r2_pcie_2.json 2> gpurun_out/r2_pcie_2.err
cat gpurun_out/r2_pcie_2.json
