set -x
python profiles/kernel_lab.py --libs default,fb256,bb256,bb320,bloop,ldg --steps 30 > gpurun_out/r2_lab1.jsonl 2> gpurun_out/r2_lab1.err
ncu --set full --clock-control none --import-source on -k regex:render_ -s 4 -c 2 -o gpurun_out/r2_render_a -f python profiles/kernel_lab.py --tag prof --steps 2 --warmup 2 > gpurun_out/r2_ncu_a.log 2>&1
cat gpurun_out/r2_lab1.jsonl
