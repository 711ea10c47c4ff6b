python -m pytest tests/test_gpu_fused.py tests/test_gpu_field.py -x -q 2>&1 | tail -5
python tools/bench_field.py --modes bf16_fused --bwd 2>&1 | tail -6
python bench.py --no-cpu --steps 10 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['fwd'], d['roofline']['bwd'], d['roofline_dw']['share_of_step'])"
