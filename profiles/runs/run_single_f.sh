set -x
cd /root/repo
python bench.py > gpurun_out/r2_bench_example.json 2> gpurun_out/r2_bench_example.err; tail -c 400 gpurun_out/r2_bench_example.json
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_driver_flags.json 2> gpurun_out/r2_bench_driver_flags.err; tail -c 300 gpurun_out/r2_bench_driver_flags.json
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
