#!/usr/bin/env python
"""Round-2 evidence: turn the gpurun_out/r2*/ scratch artefacts into tracked summaries under profiles/."""
import csv
import glob
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.isfile(os.path.join(ROOT, "MEASURED_PEAKS.json")) else \
    {"hbm_gbs": 6571.9, "bf16_tflops": 1694.6, "bf16_tflops_sustained": 1456.6}


def w(name, text):
    open(os.path.join(P, name), "w").write(text if text.endswith("\n") else text + "\n")
    print("wrote profiles/" + name)


def layer_bench(src):
    d = json.load(open(src))
    L = [f"# Per-layer kernel timings, ResNet-50 shapes, batch {d['batch']} (tools/layer_bench.py, round 2)", "",
         "CUDA events, best of 5, 256 MB L2 flush between repetitions; `ours` = the variant the per-shape autotuner picked.",
         f"`of peak` = conv FLOPs / time against the measured bf16 burst peak {peaks['bf16_tflops']:.0f} TFLOP/s; cuDNN columns: the "
         "same conv through torch/cuDNN bf16 channels_last (cuDNN does NOT produce the BN statistics our forward epilogue does).", "",
         "| shape | x | fwd us | of peak | cuDNN fwd us | dgrad us | wgrad us | cuDNN dgrad+wgrad us | fwd variants (us) | dgrad variants (us) |",
         "|---|---|---|---|---|---|---|---|---|---|"]
    fwd_ns = bwd_ns = cf_ns = cb_ns = 0.0
    for r in d["rows"]:
        v = r["stage_sweep"].get("variants_us_maxdiff", {})
        fm = lambda t: ", ".join(f"{k} {x[0]}" for k, x in t.items()) if t else "-"
        L.append(f"| {r['shape']} | {r['occ']} | {r['fwd_ms'] * 1e3:.1f} | {100 * r['fwd_tflops'] / peaks['bf16_tflops']:.0f}% | "
                 f"{r['cudnn_fwd_ms'] * 1e3:.1f} | {r['dgrad_ms'] * 1e3:.1f} | {r['wgrad_ms'] * 1e3:.1f} | {r['cudnn_bwd_ms'] * 1e3:.1f} | "
                 f"{fm(v.get('fwd', {}))} | {fm(v.get('dgrad', {}))} |")
        if not r["shape"].startswith("3x224"):
            fwd_ns += r["fwd_ms"] * r["occ"]; bwd_ns += (r["dgrad_ms"] + r["wgrad_ms"]) * r["occ"]
            cf_ns += r["cudnn_fwd_ms"] * r["occ"]; cb_ns += r["cudnn_bwd_ms"] * r["occ"]
    t = d["totals"]
    L += ["", "Totals weighted by occurrences (ms): " + ", ".join(f"{k} {v:.2f}" for k, v in t.items()),
          f"", f"52 non-stem layers: ours fwd {fwd_ns:.2f} ms vs cuDNN {cf_ns:.2f} ms; ours dgrad+wgrad {bwd_ns:.2f} ms vs cuDNN {cb_ns:.2f} ms.",
          f"All 53 layers: ours fwd+dgrad+wgrad {t['fwd'] + t['dgrad'] + t['wgrad']:.2f} ms vs cuDNN fwd+bwd {t['cudnn_fwd'] + t['cudnn_bwd']:.2f} ms "
          "(mid round 2: 11.06 vs 11.60; round 1: 11.73 vs 11.59)."]
    w("layer_bench.md", "\n".join(L))


def ncu_convlong(src):
    rows = list(csv.reader(open(src)))
    hdr = rows[0]
    want = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "us"),
            ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % (active cycles)"),
            ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor mem % "),
            ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->SM MB"), ("dram__bytes_read.sum", "DRAM read MB"),
            ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
            ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"),
            ("launch__shared_mem_per_block_dynamic", "dyn smem KB"), ("sm__cycles_active.avg", "SM active cycles")]
    idx = [(hdr.index(a), b) for a, b in want if a in hdr]
    L = ["# Nsight Compute, long-K layers through every conv kernel variant (round 2)", "",
         "`ncu --set full --clock-control none --import-source on -k regex:conv_gemm|conv_wgrad python tools/ncu_target.py convlong` "
         "(256x14x14 3x3 -> 256 and 1024x14x14 1x1 -> 256, batch 256; forward with BN-stat epilogue, then dgrad, per variant).", "",
         "Reading: the tensor pipe is busy ~55-60 % of the active cycles in EVERY variant although the L2->SM traffic differs 2x "
         "(951 MB one-tile / persistent 128x128 tiles, 715 MB deep 128x256, 476 MB cta_group::2 256x256) — operand delivery from "
         "L2 is not the limiter (951 MB in 62 us is 15 TB/s).  What the variants share is shared-memory bandwidth: an SS-mode "
         "128x128x16 MMA reads 8 KB of operands in 64 cycles = the SM's 128 B/cycle, and TMA writes the same bytes in: see "
         "DESIGN.md section 8.", "",
         "| " + " | ".join(b for _, b in idx) + " |", "|" + "---|" * len(idx)]
    for r in rows[2:]:
        cells = []
        for i, b in idx:
            v = r[i]
            if b == "kernel":
                v = re.sub(r"void unnamed>::", "", v)[:60]
            elif b in ("L2->SM MB", "DRAM read MB", "dyn smem KB"):
                pass
            cells.append(v)
        L.append("| " + " | ".join(cells) + " |")
    w("ncu_convlong.md", "\n".join(L))


def copy_text(src, dst, title, pre=""):
    if not os.path.isfile(src):
        return
    body = open(src).read()
    w(dst, f"# {title}\n\n{pre}\n```\n{body.strip()}\n```")


def sanitizer():
    L = ["# compute-sanitizer passes (tools/sanitize.sh <tool> <gpu_diag group>; round 2: tools/runs/run_san.sh + 2-GPU comm memcheck)", "",
         "| tool | kernels | configuration | summary |", "|---|---|---|---|",
         "| memcheck | conv_generic | default (round 1) | ERROR SUMMARY: 0 errors |",
         "| racecheck | bn | default (round 1) | RACECHECK SUMMARY: 0 hazards displayed (0 errors, 0 warnings) |",
         "| synccheck | conv_fwd | default (round 1) | ERROR SUMMARY: 0 errors |"]
    names = {"wgrad_mem.txt": ("memcheck", "conv_wgrad (split-K red.global epilogue, 4-D TMA operands)", "default"),
             "wgrad_race.txt": ("racecheck", "conv_wgrad", "default"),
             "deep_pair_mem.txt": ("memcheck", "conv_dgrad through the deep-ring kernel", "DDL_CONV_DEEP=2 (cta_group::2 pairs)"),
             "deep_race.txt": ("racecheck", "gemm through the deep-ring kernel", "DDL_CONV_DEEP=3 (single CTA, 256-wide tiles)"),
             "persist_race.txt": ("racecheck", "conv_fwd through the persistent kernel", "DDL_CONV_PERSISTENT=2 (CL = 0)"),
             "fp8_mem.txt": ("memcheck", "fp8 quantise + kind::f8f6f4 convs + BN twins", "default"),
             "sgd_sync.txt": ("synccheck", "fused SGD / allreduce kernels (1 GPU)", "default")}
    for f, (tool, what, cfg) in names.items():
        p = os.path.join(G, "r2s", f)
        if os.path.isfile(p):
            t = open(p).read()
            m = re.findall(r"(ERROR SUMMARY: \d+ errors|RACECHECK SUMMARY: [^\n]*)", t)
            L.append(f"| {tool} | {what} | {cfg} | {m[-1] if m else 'n/a'} |")
    p = os.path.join(G, "r2c2", "comm_memcheck.log")
    if os.path.isfile(p):
        m = re.findall(r"ERROR SUMMARY: \d+ errors", open(p).read())
        L.append(f"| memcheck | tools/comm_test.py on 2 GPUs (NVLS + P2P, both wires, block skew) | torchrun launcher process | {m[-1] if m else 'n/a'} |")
    def summary(path, pat):
        if not os.path.isfile(path):
            return None
        m = re.findall(pat, open(path).read())
        return m[-1] if m else "n/a"

    what2 = "tools/comm_test.py on 2 GPUs, `--target-processes all` (every rank): NVLS + P2P, both wires, block skew"
    r = summary(os.path.join(G, "r2t", "san_mem_N2.log"), r"ERROR SUMMARY: \d+ errors")
    if r:
        L.append(f"| memcheck | {what2} | final build | {r} |")
    r = summary(os.path.join(G, "r2t", "san_race_N2.log"), r"RACECHECK SUMMARY: [^\n]*")
    if r:
        L.append(f"| racecheck | {what2} | before the fix below | {r}: every displayed hazard is the read of `block_barrier`'s shared verdict flag "
                 "against thread 0 of the NEXT barrier resetting it (fused_allreduce_sgd.cu) |")
    r = summary(os.path.join(G, "r2u", "san_race_N2.log"), r"RACECHECK SUMMARY: [^\n]*")
    if r:
        L.append(f"| racecheck | {what2} | verdict flag indexed by epoch parity | {r} |")
    L += ["", "The racecheck finding was real but benign in normal operation (the flag is 0 unless a barrier timed out; after a "
          "time-out a straggler could have read the reset flag and gone on while its neighbours returned).  Fixed by giving "
          "the two barriers of a kernel separate flags."]
    w("sanitizer.md", "\n".join(L))


def bench_lines_r2():
    """Every bench.py JSON line of round 2 (gpurun_out/r2*/bench*.json), oldest first: the progression of the build."""
    import glob

    L = ["# bench.py JSON lines of round 2 (`gpurun_out/r2*/bench*.json`, oldest first)", "",
         "`value` = whole-job images/s, device-timed, max over ranks; `sm_mhz` = median SM clock sampled during the timed steps "
         "(1965 = no cap; lower = `sw_power_cap`).  A/B pairs (pdl0/pdl1, rev0/rev1, hint0/hint2000) ran back to back on one box.", "",
         "| run | file | GPUs | img/s | ms/step | end to end img/s | SM MHz | engine_check | note |", "|---|---|---|---|---|---|---|---|---|"]
    files = sorted(glob.glob(os.path.join(G, "r2*", "bench*.json")), key=os.path.getmtime)
    for p in files:
        try:
            line = [l for l in open(p).read().strip().splitlines() if l.startswith("{")][-1]
            d = json.loads(line)
        except Exception:
            continue
        if "value" not in d:
            continue
        e2e = (d.get("e2e") or {}).get("value")
        clk = d.get("clocks") or {}
        note = "reference script" if d.get("impl") == "reference" else ("fp8 operands" if "fp8" in str(d.get("dtype")) else "")
        if (clk.get("sm_mhz") or 1e9) < 500:
            note = "clock sampler read physical GPU 0 while the run used GPU 1 (CUDA_VISIBLE_DEVICES); fixed since: UUID"
        L.append(f"| {os.path.basename(os.path.dirname(p))} | {os.path.basename(p)} | {d.get('n_gpus')} | {d['value']:,.0f} | "
                 f"{d.get('ms_per_step'):.2f} | {e2e:,.0f} | {clk.get('sm_mhz')} | {d.get('engine_check', '')} | {note} |"
                 if e2e else
                 f"| {os.path.basename(os.path.dirname(p))} | {os.path.basename(p)} | {d.get('n_gpus')} | {d['value']:,.0f} | "
                 f"{d.get('ms_per_step'):.2f} | - | {clk.get('sm_mhz')} | {d.get('engine_check', '')} | {note} |")
    w("bench_lines_r2.md", "\n".join(L))


def main():
    bench_lines_r2()
    lb = os.path.join(G, "r2m", "layer_bench.json")           # latest full run (after the epilogue work)
    if not os.path.isfile(lb):
        lb = os.path.join(G, "r2b2", "layer_bench.json")
    if os.path.isfile(lb):
        layer_bench(lb)
    nc = os.path.join(G, "r2n", "prof_convlong_raw.csv")
    if os.path.isfile(nc):
        ncu_convlong(nc)
    copy_text(os.path.join(G, "r2b", "umma_probe.log"), "umma_swizzle_probe.md",
              "tcgen05.mma operand windows: which shared-memory bytes does a 128B-swizzled K-major descriptor read?",
              "tools/umma_probe.cu on a B200: X[256][64] bf16 is written by ONE TMA box (SWIZZLE_128B); the A descriptor then starts "
              "`shift` rows in and / or spaces its 8-row groups `sbo` bytes apart.  Result: the tensor core applies the swizzle to the "
              "ABSOLUTE shared-memory address bits exactly like TMA (base_offset must stay 0) — a haloed activation tile can feed all "
              "filter taps of a window convolution through shifted descriptors (tw = 8 pixels per group, SBO = (tw + 2) * 128).")
    for mode in ("bf16", "fp8"):
        p = os.path.join(G, "r2g", f"step_launch_list_{mode}.md")
        if os.path.isfile(p):
            w("step_launch_list.md" if mode == "bf16" else "step_launch_list_fp8.md",
              f"# One training step (ResNet-50, batch 256, 1 GPU, {mode} operands): every kernel launch with its device time\n\n"
              "`ncu --metrics gpu__time_duration.sum --clock-control none` over an eager step (serialised, cold caches: compare SHARES, "
              "not absolutes; the benchmark replays the step from a CUDA graph with the weight-gradient and bucket kernels overlapped).\n\n"
              + open(p).read())
    sanitizer()
    for n in (2, 8):
        p = os.path.join(G, f"r2c" if n == 2 else "r2f", f"timeline_N{n}.log")
        if os.path.isfile(p):
            body = open(p).read()
            body = body[body.index("model="):] if "model=" in body else body
            w(f"comm_timeline_N{n}.md", f"# Per-bucket timeline of the fused allreduce+SGD kernels inside a ResNet-50 step, {n} GPUs (tools/comm_timeline.py)\n\n"
              "Eager step (so CUDA events can sit between launches); `kernel ms` includes the time a rank waits at barrier-in for the slowest peer "
              "(eager launches skew the ranks by 0.1-0.3 ms; under CUDA-graph replay the skew and the exposed tail shrink: see BASELINE.md).\n"
              "`roofline ms` = S * (1 + 1/N) bytes / 900 GB/s.\n\n" + body)
    p = os.path.join(G, "r2x", "fp8_parity.log")
    if os.path.isfile(p):
        w("fp8_parity.md", "# FP8 training mode vs bf16: loss-curve parity (tools/fp8_parity.py --steps 200 --batch 32)\n\n"
          "Same seed, same 64-batch synthetic pool, SGD momentum 0.9, lr 0.01; fp8 = e4m3 activations / weights, e5m2 gradients for the "
          "forward and data-gradient convolutions with K >= 512, quantisation fused into the BN kernels.  bf16 runs twice: its own "
          "run-to-run gap (atomics in the reductions, a chaotic first ~100 steps) is the noise floor the fp8 curve is judged against "
          "(an earlier run of this comparison, `gpurun_out/r2v/pytest_gpu.log`, had the fp8 curve 9.8 % BELOW bf16 in the transient; "
          "another, `gpurun_out/r2e2/`, within 2.7 % throughout).\n\n" + open(p).read())
    for name, title in (("equiv_b32.log", "ResNet-50, batch 32, autotuned"), ("equiv_b32_sync.log", "ResNet-50, autotune off, single-stream wgrad"),
                        ("equiv_r18.log", "ResNet-18")):
        pass
    parts = []
    for name, title in (("equiv_b32.log", "ResNet-50, batch 32 (default configuration)"),
                        ("equiv_b32_sync.log", "ResNet-50, autotune off, weight gradients on the main stream"),
                        ("equiv_r18.log", "ResNet-18, batch 32")):
        p = os.path.join(G, "r2d", name)
        if os.path.isfile(p):
            parts.append(f"## {title}\n\n```\n{open(p).read().strip()}\n```\n")
    if parts:
        w("step_reproducibility.md", "# How reproducible is ONE training step? (tools/equiv_probe.py)\n\n"
          "Three executions of the identical first step (same seed, same batch, single-rank engine) and the cosine / relative difference of their "
          "weight updates.  fp32 atomics make reduction orders run-dependent; bf16 rounding amplifies the resulting 1e-7 differences to rounding "
          "level in every layer, and a randomly initialised ResNet-50 WITHOUT zero-init-residual has a small gradient signal at step 0 (its "
          "gradient check against fp32, `gpu_diag --group model`, uses zero_init_residual for that reason and agrees to cosine >= 0.98).  This is "
          "the run-to-run noise floor against which N-rank vs 1-rank comparisons have to be read, and why averaging N noisy copies of the same "
          "gradient makes the fixed-batch benchmark's loss fall faster with more ranks (DESIGN.md section 4.2).\n\n" + "\n".join(parts))
    sass = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_excerpts.py")], capture_output=True, text=True).stdout
    if sass:
        w("sass_excerpts.md", sass)


if __name__ == "__main__":
    main()
