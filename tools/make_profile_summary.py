"""Turn the artefacts of one `tools/gpu_round.sh` visit (gpurun_out/) into profiles/<tag>_*: the raw launch list, the HBM-kernel
benchmark log and a markdown summary (bench line, launch shares, full-capture metrics per kernel).
    python tools/make_profile_summary.py r1d "title of the snapshot" """
import collections
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
tag, title = sys.argv[1], sys.argv[2]
P = os.path.join(ROOT, "profiles")
shutil.copy(os.path.join(G, "launches.csv"), os.path.join(P, f"{tag}_launches.csv"))
shutil.copy(os.path.join(G, "bench_render.log"), os.path.join(P, f"{tag}_bench_render.log"))
shutil.copy(os.path.join(G, "bench.log"), os.path.join(P, f"{tag}_bench.json"))
shutil.copy(os.path.join(G, "bench_ref.log"), os.path.join(P, f"{tag}_bench_reference.json"))

rows = [r for r in csv.reader(open(os.path.join(G, "launches.csv"))) if len(r) > 14 and r[0].isdigit()]
agg = collections.defaultdict(lambda: [0.0, 0])
for r in rows:
    agg[r[4]][0] += float(r[14])
    agg[r[4]][1] += 1
tot = sum(v[0] for v in agg.values())
b = json.loads(open(os.path.join(G, "bench.log")).read().strip().splitlines()[-1])
ref = json.loads(open(os.path.join(G, "bench_ref.log")).read().strip().splitlines()[-1])
tests = open(os.path.join(G, "pytest_gpu.log")).read().strip().splitlines()[-2]
out = [f"# {title}\n",
       f"One `gpurun` visit (`tools/gpu_round.sh`): GPU suite `{tests.strip()}`, `smoke()`, `python bench.py` (`{tag}_bench.json`), "
       f"`bench.py --impl reference` (`{tag}_bench_reference.json`), `tools/bench_render.py` (`{tag}_bench_render.log`), "
       "`ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv python bench.py --steps 2 --warmup 3 --no-cpu --no-render` "
       f"(raw list `{tag}_launches.csv`; kernels inside the graph replays are profiled node by node; cold-cache, serialised: compare shares) "
       "and `ncu --set full` captures (one per kernel family).\n",
       f"Bench line (not under ncu): **{b['value']:.0f} train rays/s**, {b['ms_per_step']:.2f} ms/step (one CUDA graph per step, "
       f"{b['gpu_launches'] // b['steps']} kernels of this library per step), e2e {b['e2e']['value']:.0f} rays/s, render (1024x1024 eval) "
       f"**{b['render']['value']:.0f} rays/s** ({b['render']['ms_per_image']:.0f} ms/image), clocks {b['clocks']['sm_mhz']:.0f}/"
       f"{b['clocks']['sm_max_mhz']:.0f} MHz {b['clocks']['reasons']}.",
       f"Fused fwd+bwd: {b['roofline']['achieved']:.0f} TFLOP/s = {100 * b['roofline']['frac']:.1f} % of the measured sustained bf16 peak "
       f"({b['roofline']['peak']:.0f}); fwd {b['roofline']['fwd']['tflops']:.0f} TFLOP/s ({100 * b['roofline']['fwd']['share_of_step']:.1f} % of the step), "
       f"bwd {b['roofline']['bwd']['tflops']:.0f} TFLOP/s ({100 * b['roofline']['bwd']['share_of_step']:.1f} %); dW GEMM {b['roofline_dw']['achieved']:.0f} GB/s = "
       f"{100 * b['roofline_dw']['frac']:.1f} % of the measured HBM peak ({100 * b['roofline_dw']['share_of_step']:.1f} % of the step).",
       f"CPU arm on the same box: {ref['value']:.0f} rays/s on {ref['cpu_baseline']['cores']} cores (oracle port, 1024-ray slices).\n",
       "| share | total us | launches | kernel |\n|---:|---:|---:|---|"]
for k, v in sorted(agg.items(), key=lambda x: -x[1][0])[:24]:
    out.append(f"| {100 * v[0] / tot:.1f}% | {v[0] / 1e3:.0f} | {v[1]} | `{k[:90]}` |")
out.append(f"\nSum of kernel time in the captured window: {tot / 1e6:.1f} ms over {len(rows)} launches.\n")


def table(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(raw)))
    if len(rr) < 3:
        return []
    idx = {c: i for i, c in enumerate(rr[0])}
    res, seen = [], collections.Counter()
    for r in rr[2:]:
        g = lambda c: r[idx[c]] if c in idx else ""
        name = g("Kernel Name")[:48]
        seen[name] += 1
        if seen[name] > 2:
            continue
        res.append((name, float(g("gpu__time_duration.sum")), rr[1][idx["gpu__time_duration.sum"]], float(g("dram__bytes_read.sum")),
                    rr[1][idx["dram__bytes_read.sum"]], float(g("dram__bytes_write.sum")), rr[1][idx["dram__bytes_write.sum"]],
                    g("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"), g("lts__throughput.avg.pct_of_peak_sustained_elapsed"),
                    g("smsp__issue_active.avg.pct_of_peak_sustained_active"), g("launch__registers_per_thread"),
                    g("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
                    g("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed")))
    return res


reps = [("fused_full.ncu-rep", "the three dominant kernels (eager replay of the bench step)")]
reps += [(f, "`" + f[7:-8] + "` at 65 536 rays (`tools/bench_render.py`)") for f in sorted(os.listdir(G)) if f.startswith("render_") and f.endswith(".ncu-rep")]
for rep, head in reps:
    path = os.path.join(G, rep)
    if not os.path.exists(path):
        continue
    out += [f"## `ncu --set full` of {head}\n",
            "| kernel | duration | dram read | dram write | tensor pipe active % | lts throughput % | issue active % | regs | smem wavefronts LSU % | smem wavefronts tensor % |",
            "|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|"]
    for r in table(path):
        out.append(f"| `{r[0]}` | {r[1]:.3f} {r[2]} | {r[3]:.2f} {r[4]} | {r[5]:.2f} {r[6]} | {float(r[7] or 0):.1f} | {float(r[8] or 0):.1f} | {float(r[9] or 0):.1f} | {r[10]} | {float(r[11] or 0):.1f} | {float(r[12] or 0):.1f} |")
    out.append("")
# per-launch DRAM traffic of the dominant kernels for bench.py's roofline.traffic (read by bench.py from profiles/traffic.json)
ff = table(os.path.join(G, "fused_full.ncu-rep")) if os.path.exists(os.path.join(G, "fused_full.ncu-rep")) else []
unit = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
fused = [(r[3] * unit[r[4]] + r[5] * unit[r[6]]) for r in ff if "fused_fwd" in r[0] or "fused_bwd" in r[0]]
dw = [(r[3] * unit[r[4]] + r[5] * unit[r[6]]) for r in ff if "gemm_tn_blocked" in r[0]]
if fused:
    json.dump({"source": f"profiles/{tag}_summary.md: ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, mean over the captured launches",
               "fused_fwd_bwd_bytes_per_launch": sum(fused) / len(fused), "fused_launches": len(fused),
               "dw_gemm_bytes_per_launch": (sum(dw) / len(dw)) if dw else None, "dw_launches": len(dw)},
              open(os.path.join(P, "traffic.json"), "w"), indent=1)
if os.path.exists(os.path.join(G, "bench_field.log")):
    out.append("## Fused field kernels alone, 1 M samples, CUDA events (`tools/bench_field.py --bwd`; bwd = input-gradient chain + dW GEMM + head gradients)\n\n```")
    out += [l.rstrip() for l in open(os.path.join(G, "bench_field.log")).read().splitlines()]
    out.append("```\n")
out.append("## HBM-bound SIMT kernels, CUDA events (`tools/bench_render.py`)\n\n```")
out += [l.rstrip() for l in open(os.path.join(G, "bench_render.log")).read().splitlines() if not l.startswith("{")]
out.append("```")
open(os.path.join(P, f"{tag}_summary.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
