TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
B="bench.py --steps 40 --warmup 5 --no-cpu --no-e2e"
echo "== torchrun x2 (NUMA bind off)"; SURFEL_BENCH_NO_NUMA=1 $TR --nproc-per-node 2 --master-port 29512 $B --gpus 2 2>/dev/null | grep '^{' | python profiles/brief.py
echo "== torchrun x1"; $TR --nproc-per-node 1 --master-port 29513 $B --gpus 1 2>/dev/null | grep '^{' | python profiles/brief.py
echo "== two plain processes"; CUDA_VISIBLE_DEVICES=0 python $B > /tmp/p0.json 2>/dev/null & CUDA_VISIBLE_DEVICES=1 python $B > /tmp/p1.json 2>/dev/null; wait; grep '^{' /tmp/p0.json | python profiles/brief.py; grep '^{' /tmp/p1.json | python profiles/brief.py
echo "== torchrun x2 OMP unset"; $TR --nproc-per-node 2 --master-port 29514 $B --gpus 2 2>/dev/null | grep '^{' | python profiles/brief.py
nproc; lscpu | grep -i "numa\|model name" | head -6; nvidia-smi topo -m | head -8
