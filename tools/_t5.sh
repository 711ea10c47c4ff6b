python -m pytest tests/test_gpu_static.py -x -q 2>&1 | tail -15
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --no-cpu --steps 20 > gpurun_out/bench_graph.log 2> gpurun_out/bench_graph.err; echo "bench exit $?"; tail -3 gpurun_out/bench_graph.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_graph.log').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'], d['roofline']['fwd'], d['roofline']['bwd'], d['roofline_dw']['share_of_step'], d['config']['kept_samples_per_step_per_gpu'])"
python bench.py --no-cpu --steps 20 --no-graph 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('eager', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'])"
