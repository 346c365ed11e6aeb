set -x
cd /root/repo
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$R --master-port 29533 profiles/band_probe.py > gpurun_out/r2_band_probe_default.json 2> gpurun_out/r2_band_probe.err
TORCH_NCCL_AVOID_RECORD_STREAMS=1 $R --master-port 29535 profiles/band_probe.py > gpurun_out/r2_band_probe_avoid.json 2>> gpurun_out/r2_band_probe.err
PYTORCH_CUDA_ALLOC_CONF=expandable_segments:True $R --master-port 29537 profiles/band_probe.py > gpurun_out/r2_band_probe_expandable.json 2>> gpurun_out/r2_band_probe.err
tail -c 1500 gpurun_out/r2_band_probe_default.json; tail -5 gpurun_out/r2_band_probe.err
