"""profiles/scaling_<tag>.txt from the bench lines of the 1/2/4/8-GPU runs kept in gpurun_out/
(bench_final.json, bench_n2/n4/n8.json: config 3/4; bench_c5b.json, bench_c5_n2/n4/n8.json: config 5).

  python tools/make_scaling.py [r02]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = os.path.join(ROOT, "gpurun_out")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r02"


def load(f):
    return json.loads(open(os.path.join(R, f)).read().strip().splitlines()[-1])


out = ["# Multi-GPU scaling, round 2 (one process per GPU, torchrun, NCCL all-reduce of [gradient | 16 scalars]; device time = max over ranks)",
       "# produced by: python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N [--config 5 --steps 10] --no-cpu-baseline",
       "",
       "## BASELINE config 3/4: NT=60000 images sharded over N GPUs (strong scaling), N=196 sites, maxm=120, Npass=4; 20 class-L bond updates (10 rightwards + 10 leftwards)",
       "| GPUs | bond-updates/s | ms/bond | speed-up | proj | grad | fat | svd (replicated) | shift | rest (NCCL, launch gaps, syncs) | second full sweep, bond-updates/s |",
       "|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|"]
base = None
svd1 = svd8 = None
for n, f in [(1, "bench_final.json"), (2, "bench_n2.json"), (4, "bench_n4.json"), (8, "bench_n8.json")]:
    d = load(f)
    ph = d["roofline"]["phase_ms_per_step"]
    pk = list(ph.keys())
    v = d["value"]
    base = base or v
    rest = d["ms_per_step"] - sum(ph.values())
    sa = d.get("value_sweep_avg")
    sa = f"{sa['value']:.1f}" if sa else "-"
    if n == 1:
        svd1 = ph[pk[3]] / d["ms_per_step"]
    if n == 8:
        svd8 = ph[pk[3]] / d["ms_per_step"]
    out.append(f"| {n} | {v:.1f} | {d['ms_per_step']:.3f} | {v / base:.2f}x | {ph[pk[0]]:.3f} | {ph[pk[1]]:.3f} | {ph[pk[2]]:.3f} | "
               f"{ph[pk[3]]:.3f} | {ph[pk[4]]:.3f} | {rest:.3f} | {sa} |")
out += ["",
        f"The truncated SVD (2.4-2.5 ms, one 16-CTA cluster) is replicated on every rank and does not shrink: it is {100 * svd1:.0f} % of the "
        f"step on one GPU and {100 * svd8:.0f} % on eight.",
        "",
        "## BASELINE config 5: synthetic 1e6 images sharded over N GPUs, maxm=minm=300, window of bonds 2..6 (class L, C, C, R, R) of an "
        "8-site chain, rightwards then leftwards (10 bond updates)",
        "| GPUs | bond-updates/s | ms/bond | speed-up | images x bonds / s | proj | grad | fat | svd | shift |",
        "|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|"]
base = None
for n, f in [(1, "bench_c5b.json"), (2, "bench_c5_n2.json"), (4, "bench_c5_n4.json"), (8, "bench_c5_n8.json")]:
    d = load(f)
    ph = d["roofline"]["phase_ms_per_step"]
    pk = list(ph.keys())
    v = d["value"]
    base = base or v
    out.append(f"| {n} | {v:.3f} | {d['ms_per_step']:.1f} | {v / base:.2f}x | {v * 1e6:.3g} | {ph[pk[0]]:.1f} | {ph[pk[1]]:.1f} | {ph[pk[2]]:.1f} | "
               f"{ph[pk[3]]:.1f} | {ph[pk[4]]:.1f} |")
out += ["",
        "At this size (link dimension 300 > 128) the projection runs on the FP64 mma.sync kernels (the tcgen05 kernel keeps K <= 128 resident), the",
        "class-C bonds (label index on the bond tensor, 10x the projection work) dominate the window, and the serial SVD (600 x 600 and 6000 x 600:",
        "27 ms) is 1.6 % of the one-GPU step and 11 % of the eight-GPU step: 7.1x on 8 GPUs.",
        "(Config 5 was measured a few commits before the config-3/4 table; nothing on its code path changed since.)"]
open(os.path.join(ROOT, "profiles", f"scaling_{TAG}.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out[:14]))
