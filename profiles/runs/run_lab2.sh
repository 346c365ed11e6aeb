set -x
python profiles/kernel_lab.py --libs default,bb192,bb224,bb256,bs2,bs0,fb192,fb256,fb256o6 --steps 30 > gpurun_out/r2_lab2.jsonl 2> gpurun_out/r2_lab2.err
SURFEL_LIB=$PWD/2d-gaussian-splatting_b200/lib/variants/bb256.so ncu --set full --clock-control none --import-source on -k regex:render_ -s 4 -c 2 -o gpurun_out/r2_render_a -f python profiles/kernel_lab.py --tag prof --steps 2 --warmup 2 > gpurun_out/r2_ncu_a.log 2>&1
ls -la gpurun_out
cat gpurun_out/r2_lab2.jsonl
