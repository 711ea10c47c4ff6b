python -m pytest tests/test_gpu_static.py tests/test_gpu_fused.py tests/test_gpu_render.py -x -q 2>&1 | tail -2
python tools/profile_step.py 2>&1 | grep -E "heads_dw_blocked|class_grad|total device"
python tools/profile_step.py --graph 2>&1 | grep -E "heads_dw_blocked|class_grad|total device"
