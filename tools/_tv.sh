python -m pytest tests/test_gpu_fused.py tests/test_gpu_static.py tests/test_gpu_render.py tests/test_gpu_field.py -x -q 2>&1 | tail -2
for w in 4 3 4; do
  EONERF_EXTRA_NVCC_FLAGS=-DEONERF_FWD_RING=$w python -m eonerf_code_b200.build --force > /dev/null 2>&1
  echo "FWD_RING=$w"
  python tools/bench_field.py --modes bf16_fused 2>&1 | grep -E "fwd" | tr '\n' ';'; echo
done
