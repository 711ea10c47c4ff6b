#!/bin/bash
# One GPU box visit: GPU suite, smoke, bench (both arms), kernel micro-benchmarks, ncu launch list and full captures of the
# dominant kernels.  Everything lands in gpurun_out/ (summaries are copied to profiles/ by tools/make_profile_summary.py).
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -4 gpurun_out/smoke.log
python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; cat gpurun_out/bench_ref.log
python tools/bench_render.py > gpurun_out/bench_render.log 2>&1; head -9 gpurun_out/bench_render.log
python tools/bench_field.py --modes bf16_fused --bwd > gpurun_out/bench_field.log 2>&1; cat gpurun_out/bench_field.log
# launch list of the default (CUDA graph) bench: kernels inside graph replays are profiled node by node
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-render --no-extra > gpurun_out/ncu_list.log 2>&1; echo "ncu list exit $?"
# full captures without source import: the visit's gpurun_out/ must stay under 64 MiB to be copied back
ncu --set full --clock-control none -k regex:'fused_fwd_kernel|fused_bwd_kernel|gemm_tn_blocked_kernel' -s 12 -c 6 \
    -o gpurun_out/fused_full -f python bench.py --steps 1 --warmup 3 --no-cpu --no-render --no-extra --no-graph > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?"
for k in sample_onepass_kernel composite_fwd_kernel composite_bwd_kernel shadow_fwd_kernel accumulate_fwd_vec_kernel; do
  ncu --set full --clock-control none -k regex:$k -s 5 -c 1 -o gpurun_out/render_$k -f python tools/bench_render.py > gpurun_out/ncu_render_$k.log 2>&1; echo "ncu $k exit $?"
done
ls -la gpurun_out | tail -30; du -sh gpurun_out
