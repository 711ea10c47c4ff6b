EONERF_EXTRA_NVCC_FLAGS=-DEONERF_TIMING python -m eonerf_code_b200.build --force > /dev/null 2>&1
for dbg in 0 2 4 6; do echo "DBG=$dbg"; EONERF_FUSED_DBG=$dbg python tools/fused_timing.py 2>&1 | grep "keep=1"; done
python tools/fused_timing_bwd.py 2>&1 | tail -2
