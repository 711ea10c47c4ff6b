python -m pytest tests/test_gpu_static.py tests/test_gpu_fused.py -x -q 2>&1 | tail -2
for u in 8 4; do
  EONERF_EXTRA_NVCC_FLAGS=-DEONERF_HEADS_UNROLL=$u python -m eonerf_code_b200.build --force > /dev/null 2>&1
  echo "HEADS_UNROLL=$u"; python tools/profile_step.py 2>&1 | grep -E "heads_dw_blocked|class_grad|total device"
done
python bench.py --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['value'], d['roofline']['traffic'], d['roofline_dw']['traffic'], d['roofline_dw']['algorithmic_bytes_per_launch'], d['cpu_baseline'])"
