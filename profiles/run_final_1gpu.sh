# Final evidence run of a round on ONE B200: GPU test suite, the bench line, the reference arm (plain and under
# torchrun x2, where rank 0 alone works), the ncu launch list of a bench run, and one `--set full` capture of a
# step's kernels.  Everything lands in gpurun_out/; copy what should be judged into profiles/.
set -x
cd /root/repo
rm -f gpurun_out/parity_stats.jsonl
python -m pytest tests -q -m gpu -rf 2>&1 | tail -40 > gpurun_out/r2_pytest_final.txt; tail -4 gpurun_out/r2_pytest_final.txt
python bench.py > gpurun_out/r2_bench_example.json 2> gpurun_out/r2_bench_example.err; tail -c 600 gpurun_out/r2_bench_example.json
python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2_bench_driver_flags.json 2> gpurun_out/r2_bench_driver_flags.err; tail -c 300 gpurun_out/r2_bench_driver_flags.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2_bench_reference_arm_torchrun2.json 2> gpurun_out/r2_bench_reference_arm_torchrun2.err
tail -c 700 gpurun_out/r2_bench_reference_arm.json; tail -c 700 gpurun_out/r2_bench_reference_arm_torchrun2.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'preprocess|tile_|render' --launch-skip 36 --launch-count 9 -f -o gpurun_out/r2_all_kernels python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2_ncu_full.log 2>&1
tail -3 gpurun_out/r2_ncu_full.log; ls -la gpurun_out/r2_all_kernels.ncu-rep
for tool in memcheck racecheck initcheck synccheck; do
  timeout 240 compute-sanitizer --tool $tool --print-limit 5 python profiles/sanitizer_workload.py 2>&1 | tail -4 > gpurun_out/r2_sanitizer_$tool.txt
  tail -2 gpurun_out/r2_sanitizer_$tool.txt
done
