#!/usr/bin/env python
"""Turn the scratch artefacts in gpurun_out/ into the tracked summaries under profiles/."""
import csv
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)
peaks = {"hbm_gbs": 6571.9, "bf16_tflops": 1694.6, "bf16_tflops_sustained": 1456.6}
try:
    peaks.update(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))))
except Exception:
    pass


def layer_bench():
    p = os.path.join(G, "layer_bench.json")
    if not os.path.isfile(p):
        return
    d = json.load(open(p))
    L = ["# Per-layer kernel timings, ResNet-50 shapes, batch %d (tools/layer_bench.py)" % d["batch"], "",
         "CUDA events, best of 5, 256 MB L2 flush between repetitions.  `TF/s` = conv FLOPs / time (of measured bf16 peak "
         f"{peaks['bf16_tflops']:.0f} TFLOP/s burst), `GB/s` = (input + output activation bytes) / time (of measured copy "
         f"{peaks['hbm_gbs']:.0f} GB/s).  cuDNN columns: the same conv through torch/cuDNN bf16 channels_last, for orientation only.", "",
         "| shape | x | fwd ms | fwd TF/s | of peak | fwd GB/s | of copy | BN fwd ms | dgrad ms | wgrad ms | BN bwd ms (mask z / mask x) | cuDNN fwd | cuDNN bwd |",
         "|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
    for r in d["rows"]:
        L.append(f"| {r['shape']} | {r['occ']} | {r['fwd_ms']:.3f} | {r['fwd_tflops']:.0f} | {100 * r['fwd_tflops'] / peaks['bf16_tflops']:.0f}% | "
                 f"{r['fwd_gbs']:.0f} | {100 * r['fwd_gbs'] / peaks['hbm_gbs']:.0f}% | {r['bn_ms']:.3f} | {r['dgrad_ms']:.3f} | {r['wgrad_ms']:.3f} | "
                 f"{r['bnb_ms']:.3f} / {r.get('bnb_maskx_ms', 0):.3f} | {r['cudnn_fwd_ms']:.3f} | {r['cudnn_bwd_ms']:.3f} |")
    t = d["totals"]
    L += ["", "Totals weighted by occurrences (ms): " + ", ".join(f"{k} {v:.2f}" for k, v in t.items())]
    open(os.path.join(P, "layer_bench.md"), "w").write("\n".join(L) + "\n")


def launches():
    p = os.path.join(G, "launches.csv")
    if not os.path.isfile(p):
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "parse_launches.py"), p], capture_output=True, text=True).stdout
    open(os.path.join(P, "step_launch_list.md"), "w").write(
        "# One training step (ResNet-50, batch 256, 1 GPU): every kernel launch with its device time\n\n"
        "`ncu --metrics gpu__time_duration.sum --clock-control none` (serialised, cold caches: compare SHARES, not absolutes).\n\n" + out)


def ncu_reports():
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
            "launch__block_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__shared_mem_per_block_dynamic"]
    L = ["# Nsight Compute captures (`ncu --set full --clock-control none --import-source on`, 1 GPU, tools/ncu_target.py)", "",
         "Raw reports are scratch (gpurun_out/*.ncu-rep); the table keeps the roofline-relevant counters.", ""]
    for rep in sorted(glob.glob(os.path.join(G, "*.ncu-rep"))):
        r = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True)
        rows = list(csv.reader(r.stdout.splitlines()))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        idx = [hdr.index(w) for w in want if w in hdr]
        L += [f"## {os.path.basename(rep)}", "", "| " + " | ".join(hdr[i].replace('launch__', '').replace('.avg.pct_of_peak_sustained', ' %') for i in idx) + " |",
              "|" + "---|" * len(idx)]
        for row in rows[2:]:
            L.append("| " + " | ".join((row[i][:70] + (" " + units[i] if units[i] and i != idx[0] else "")) for i in idx) + " |")
        L.append("")
    open(os.path.join(P, "ncu_summary.md"), "w").write("\n".join(L) + "\n")


def sanitizer():
    logs = sorted(glob.glob(os.path.join(G, "sanitize_*.log")))
    if not logs:
        return
    L = ["# compute-sanitizer passes (tools/sanitize.sh <tool> <gpu_diag group>, 1 GPU)", "",
         "| tool | kernel group (tools/gpu_diag.py) | summary |", "|---|---|---|"]
    for f in logs:
        name = os.path.basename(f)[len("sanitize_"):-len(".log")]
        tool, group = name.split("_", 1)
        txt = open(f, errors="replace").read()
        summ = [l.strip("= ").strip() for l in txt.splitlines() if "SUMMARY" in l]
        rc = [l for l in txt.splitlines() if l.startswith("compute-sanitizer")]
        L.append(f"| {tool} | {group} | {'; '.join(summ) or 'n/a'} ({rc[-1].split()[-1] if rc else 'rc=?'}) |")
    open(os.path.join(P, "sanitizer.md"), "w").write("\n".join(L) + "\n")


def comm():
    for p in sorted(glob.glob(os.path.join(G, "comm_sweep_N*.json"))):
        d = json.load(open(p))
        n = d["world"]
        L = [f"# In-place allreduce sweep, {n} x B200 (tools/comm_test.py)", "",
             "Device-timed (CUDA events), max over ranks; busbw = algbw * 2(N-1)/N.  Reference points from the profiling guide: "
             "NCCL 8-rank bus bandwidth 725 GB/s at 1 GiB, peer copy 770 GB/s per direction (900 nominal).",
             f"Multicast (NVLS) available: {d['multicast']}; fused-engine correctness checks: {'all ok' if d['engine_checks_ok'] else 'FAILED'}.", "",
             "| bytes | dtype | impl | blocks | time us | algbw GB/s | busbw GB/s | busbw / 770 |", "|---|---|---|---|---|---|---|---|"]
        best = {}
        for r in d["results"]:
            k = (r["bytes"], r["dtype"], r["impl"])
            if k not in best or r["ms"] < best[k]["ms"]:
                best[k] = r
        for (b, dt, impl), r in sorted(best.items()):
            L.append(f"| {b} | {dt} | {impl} | {r['blocks']} | {r['ms'] * 1e3:.1f} | {r['algbw_gbs']:.1f} | {r['busbw_gbs']:.1f} | {r['busbw_gbs'] / 770:.2f} |")
        open(os.path.join(P, f"comm_sweep_N{n}.md"), "w").write("\n".join(L) + "\n")


def bench_lines():
    L = ["# bench.py JSON lines collected this round (gpurun_out/*.json, oldest first: the progression of the build)", ""]
    for p in sorted(glob.glob(os.path.join(G, "*.json")), key=os.path.getmtime):
        if os.path.basename(p).startswith((".", "layer_bench", "comm_sweep")):
            continue
        try:
            txt = open(p).read().strip().splitlines()
            line = [l for l in txt if l.startswith("{")][-1]
            d = json.loads(line)
        except Exception:
            continue
        if "value" not in d:
            continue
        e2e = d.get("e2e") or {}
        L.append(f"* `{os.path.basename(p)}`: impl={d.get('impl')} n_gpus={d.get('n_gpus')} value={d['value']:.1f} {d.get('unit')} "
                 f"ms/step={d.get('ms_per_step'):.2f} e2e={e2e.get('value', float('nan')):.1f} launches={d.get('gpu_launches')} "
                 f"graph={(d.get('config') or {}).get('cuda_graph')} clocks={d.get('clocks')}")
    open(os.path.join(P, "bench_lines.md"), "w").write("\n".join(L) + "\n")


if __name__ == "__main__":
    layer_bench()
    launches()
    ncu_reports()
    sanitizer()
    comm()
    bench_lines()
    print(sorted(os.listdir(P)))
