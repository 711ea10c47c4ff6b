#!/bin/bash
# One GPU box visit: GPU suite, smoke, bench (both arms), ncu launch list and full captures of the dominant kernels.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -4 gpurun_out/smoke.log
python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; cat gpurun_out/bench_ref.log
python tools/bench_render.py > gpurun_out/bench_render.log 2>&1; head -8 gpurun_out/bench_render.log
# launch list of the default (CUDA graph) bench: kernels inside graph replays are profiled node by node
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-render > gpurun_out/ncu_list.log 2>&1; echo "ncu list exit $?"
# full captures without source import: the visit's gpurun_out/ must stay under 64 MiB to be copied back
ncu --set full --clock-control none -k regex:'fused_fwd_kernel|fused_bwd_kernel|gemm_tn_blocked_kernel' -s 12 -c 6 \
    -o gpurun_out/fused_full -f python bench.py --steps 1 --warmup 3 --no-cpu --no-render --no-graph > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?"
for k in composite_fwd_kernel composite_bwd_kernel shadow_fwd_kernel shadow_bwd_kernel weights_fwd_kernel; do
  ncu --set full --clock-control none -k $k -s 5 -c 1 -o gpurun_out/render_$k -f python tools/bench_render.py > gpurun_out/ncu_render_$k.log 2>&1; echo "ncu $k exit $?"
done
ls -la gpurun_out; du -sh gpurun_out
