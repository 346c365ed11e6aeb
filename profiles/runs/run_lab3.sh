set -x
python profiles/kernel_lab.py --libs default,bb192,bb256,bs1,bs3,fb256,fb256o6 --steps 30 > gpurun_out/r2_lab3.jsonl 2> gpurun_out/r2_lab3.err
ncu --set full --clock-control none --import-source on -k regex:render_ -s 4 -c 2 -o gpurun_out/r2_render_b -f python profiles/kernel_lab.py --tag prof --steps 2 --warmup 2 > gpurun_out/r2_ncu_b.log 2>&1
ls -la gpurun_out
cat gpurun_out/r2_lab3.jsonl
