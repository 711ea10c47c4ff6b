python -m pytest tests/test_gpu_fused.py -x -q 2>&1 | tail -2
python tools/bench_field.py --modes bf16_fused 2>&1 | tail -4
ncu --set full --clock-control none --import-source on -k regex:fused_fwd_kernel -s 2 -c 8 -o gpurun_out/fwd_only -f python tools/bench_field.py --modes bf16_fused > gpurun_out/ncu_fwd_only.log 2>&1; echo "ncu exit $?"
