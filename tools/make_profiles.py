"""Summarise ncu outputs brought back in gpurun_out/ into tracked files under profiles/.

  python tools/make_profiles.py r01

Inputs (produced on the GPU box, see profiles/README.md for the exact commands):
  gpurun_out/launches_<tag>.csv      ncu --metrics gpu__time_duration.sum launch list of bench.py
  gpurun_out/prof_<kernel>.ncu-rep   ncu --set full capture of one launch of each hot kernel
"""
import collections
import csv
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out = os.path.join(ROOT, "profiles")
os.makedirs(out, exist_ok=True)

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_op_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__ops_path_tensor_op_utcimma_src_int8_sparsity_off.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed_op_tma_ld.sum", "smsp__sass_inst_executed_op_utcmma.sum",
        "smsp__sass_inst_executed_op_tmem_ldt.sum", "sm__cycles_active.avg", "sm__cycles_elapsed.avg.per_second",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed.sum"]


def launch_list(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        k = row["Kernel Name"].split("(")[0].replace("void ", "").replace("tnml::", "")
        agg.setdefault(k, []).append(v)
    tot = sum(sum(v) for v in agg.values())
    rows = [dict(kernel=k, launches=len(v), total_us=round(sum(v), 1), avg_us=round(sum(v) / len(v), 2),
                 share=round(sum(v) / tot, 4)) for k, v in agg.items()]
    rows.sort(key=lambda r: -r["total_us"])
    return rows, tot


def rep_summary(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        return None
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    res = {"kernel": d.get("Kernel Name", ("", ""))[0][:120]}
    for k in KEYS:
        if k in d and d[k][0] not in ("", "n/a"):
            res[k] = f"{d[k][0]} {d[k][1]}".strip()
    st = [(h.split("stalled_")[-1], float(v[0].replace(",", ""))) for h, v in d.items()
          if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h and v[0] not in ("", "n/a")]
    tot = sum(v for _, v in st) or 1.0
    res["stall_top"] = {h: round(v / tot, 3) for h, v in sorted(st, key=lambda x: -x[1])[:6]}
    return res


summary = {"tag": tag}
ll = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
md = [f"# ncu summaries, {tag}\n"]
if os.path.exists(ll):
    rows, tot = launch_list(ll)
    summary["launch_list"] = rows
    md.append(f"## Launch list of `bench.py` (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: "
              f"compare shares)\n\nTotal {tot / 1e3:.2f} ms over {sum(r['launches'] for r in rows)} launches.\n")
    md.append("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|")
    for r in rows:
        md.append(f"| `{r['kernel']}` | {r['launches']} | {r['total_us']} | {r['avg_us']} | {r['share']:.3f} |")
    md.append("")
reps = sorted(glob.glob(os.path.join(ROOT, "gpurun_out", f"prof_{tag}_*.ncu-rep")))
summary["kernels"] = {}
for rp in reps:
    name = os.path.basename(rp)[len(f"prof_{tag}_"):-len(".ncu-rep")]
    s = rep_summary(rp)
    if not s:
        continue
    summary["kernels"][name] = s
    md.append(f"## `{name}` (ncu --set full, one launch)\n")
    for k, v in s.items():
        md.append(f"* {k}: {v}")
    md.append("")
# what bench.py's roofline.traffic reads: dram bytes of one launch of the dominant kernel, the NT it was
# captured at and the commit of the build
def _bytes(x):
    v, u = x.split()[0], x.split()[1] if len(x.split()) > 1 else "byte"
    f = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    return float(v.replace(",", "")) * f


try:
    summary["commit"] = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True,
                                       cwd=ROOT).stdout.strip()
except Exception:
    summary["commit"] = "?"
if "oz_gemm" in summary["kernels"]:
    k = summary["kernels"]["oz_gemm"]
    try:
        k["dram_bytes"] = _bytes(k["dram__bytes_read.sum"]) + _bytes(k["dram__bytes_write.sum"])
        k["NT"] = 30000
    except Exception:
        pass
json.dump(summary, open(os.path.join(out, f"{tag}_ncu_summary.json"), "w"), indent=1)
open(os.path.join(out, f"{tag}_ncu_summary.md"), "w").write("\n".join(md) + "\n")
print("wrote", os.path.join(out, f"{tag}_ncu_summary.md"))
