python tools/bench_field.py --modes bf16_fused --bwd 2>&1 | tail -12
EONERF_FUSED_MODE=1 python tools/bench_field.py --modes bf16_fused 2>&1 | tail -5
EONERF_EXTRA_NVCC_FLAGS=-DEONERF_TIMING python -m eonerf_code_b200.build --force > /dev/null 2>&1
python tools/fused_timing.py 2>&1 | tail -6
EONERF_FUSED_MODE=1 python tools/fused_timing.py 2>&1 | tail -6
