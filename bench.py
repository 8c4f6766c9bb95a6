#!/usr/bin/env python
"""bench.py -- bond-updates/sec of the fixedL hot path on B200 (BASELINE.json metric).

One "step" = one iteration of the mldmrg loop body (fixedL.cc:478-563):
setBond + cgrad(Npass=4) + svd + quadcost(newB) + shiftE, on BASELINE config 3
shapes: synthetic MNIST-shaped 14x14 images (N=196 sites, d=2, 10 labels),
NT=60000 images in total, maxm=120 (minm=60), cutoff 1e-10, float64.  With
--gpus N the 60000 images are sharded over N ranks like ParallelDo bounds
(BASELINE config 4, strong scaling) and the gradient is all-reduced by NCCL.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nt", type=int, default=60000, help="total training images (config 3: 60000)")
    ap.add_argument("--maxm", type=int, default=120)
    ap.add_argument("--npass", type=int, default=4)
    ap.add_argument("--first-bond", type=int, default=10)
    ap.add_argument("--cpu-sample", type=int, default=1024, help="images in the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cg-reuse-forward", type=int, default=0,
                    help="1: linear update of the forward outputs between CG passes (tnml_set_option); "
                         "0 (default): literal recompute like fixedL.cc:412-421")
    return ap.parse_args()


# ---------------------------------------------------------------------------
def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons during the timed region: NVML (nvidia_ml_py, ~1 ms per sample),
    falling back to nvidia-smi (the recipe's clocks line)."""

    def __init__(self, gpu_index, uuid=None):
        self.rows = []          # (sm_mhz, sm_max_mhz, [reasons])
        self.stop = False
        self.idx = gpu_index
        self.uuid = uuid
        self.t = None
        self.how = None

    def _run_nvml(self):
        import pynvml as N
        N.nvmlInit()
        try:
            h = N.nvmlDeviceGetHandleByUUID(self.uuid) if self.uuid else N.nvmlDeviceGetHandleByIndex(self.idx)
        except Exception:
            h = N.nvmlDeviceGetHandleByIndex(self.idx)
        mx = float(N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM))
        bits = {"hw_slowdown": N.nvmlClocksEventReasonHwSlowdown if hasattr(N, "nvmlClocksEventReasonHwSlowdown")
                else N.nvmlClocksThrottleReasonHwSlowdown,
                "hw_thermal_slowdown": getattr(N, "nvmlClocksEventReasonHwThermalSlowdown",
                                               getattr(N, "nvmlClocksThrottleReasonHwThermalSlowdown", 0)),
                "sw_thermal_slowdown": getattr(N, "nvmlClocksEventReasonSwThermalSlowdown",
                                               getattr(N, "nvmlClocksThrottleReasonSwThermalSlowdown", 0)),
                "sw_power_cap": getattr(N, "nvmlClocksEventReasonSwPowerCap",
                                        getattr(N, "nvmlClocksThrottleReasonSwPowerCap", 0))}
        get = getattr(N, "nvmlDeviceGetCurrentClocksEventReasons", None) or N.nvmlDeviceGetCurrentClocksThrottleReasons
        self.how = "nvml"
        while not self.stop:
            sm = float(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM))
            r = int(get(h))
            self.rows.append((sm, mx, [k for k, b in bits.items() if b and (r & b)]))
            time.sleep(0.005)
        N.nvmlShutdown()

    def _run_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        self.how = "nvidia-smi"
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 6:
                    self.rows.append((float(f[0]), float(f[1]),
                                      [n for i, n in enumerate(names) if f[2 + i].lower().startswith("active")]))
            except Exception:
                pass
            time.sleep(0.05)

    def _run(self):
        try:
            self._run_nvml()
        except Exception:
            self._run_smi()

    def start(self):
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def finish(self):
        self.stop = True
        if self.t:
            self.t.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(r[0] for r in self.rows)
        reasons = sorted({x for r in self.rows for x in r[2]})
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.rows[0][1], "reasons": reasons,
                "samples": len(self.rows), "how": self.how}


def build_workload(args, rank, world):
    """Synthetic config-3 shaped inputs for this rank's shard."""
    from tnml_b200 import data, fixedl
    NTg = args.nt
    b0, b1 = fixedl.bounds(world, NTg)[rank]
    pix, labels = data.synthetic_digits(b1 - b0, 14, seed=20260925, first=b0)
    feat = data.phi(pix)                      # [NT, 196, 2] = TState::data
    W = data.random_mps(196, 2, args.maxm, seed=3)
    return feat, labels, W, NTg, b0


def cpu_baseline(args, steps=1):
    """The reference's own CPU formulation -- dense t.v per image (fixedL.cc:183-185),
    per-image P = B*t.v / dP*dag(t.v) loops (349-445, 280-344), ParallelDo threads
    (paralleldo.h) -- as restated in oracle/fixedl_ref_cpu.cpp, timed on this box's
    host cores on a bounded sample of the workload and scaled linearly in NT."""
    import tempfile
    from oracle import fixedl_oracle as O           # checker / baseline only
    from oracle import cpu_ref
    from tnml_b200 import data
    ns = args.cpu_sample
    pix, labels = data.synthetic_digits(ns, 14, seed=20260925)
    feat = O.features(pix)
    W = data.random_mps(196, 2, args.maxm, seed=3)
    ts = O.TrainStates(feat, labels.astype(np.int64), 1)
    ts.init(W)
    b = args.first_bond
    for bb in range(1, b):
        ts.set_bond(bb)
        ts.shiftE(W, bb, "Fromleft")
    ts.set_bond(b)
    B = O.form_bond(W[b], W[b + 1])
    nthread = min(16, os.cpu_count() or 1)          # paralleldo.h:55-56 caps the reference at 16
    with tempfile.TemporaryDirectory() as td:
        prob, res = os.path.join(td, "p.bin"), os.path.join(td, "r.bin")
        cpu_ref.write_problem(prob, *cpu_ref.problem_from_oracle(ts, B), Npass=args.npass)
        out = cpu_ref.run(prob, res, nthread, max(2, steps), B.shape)   # best of >=2 (first touch of 5 GB is slow)
    t = float(out["t_setbond"] + out["t_cgrad"] + out["t_quadcost"])
    bonds_per_s = 1.0 / (t * args.nt / ns)
    return {"value": bonds_per_s, "unit": "bond-updates/sec", "cores": int(nthread), "kind": "port",
            "sample": f"{ns} images, one bond update at ml=mr={args.maxm} (best of {max(2, steps)}): setBond {out['t_setbond']:.2f} s + "
                      f"cgrad {out['t_cgrad']:.2f} s + quadcost {out['t_quadcost']:.2f} s with {nthread} std::async "
                      f"threads (C++ -O3 literal dense-t.v port of fixedL.cc; svd/shiftE not included), scaled "
                      f"linearly to NT={args.nt}",
            "host_cpus": os.cpu_count()}, t


def config_dict(args, world):
    return {"workload": "BASELINE config 3/4: synthetic 14x14 MNIST-shaped, N=196 sites d=2 NL=10, "
                        f"NT={args.nt} images total, maxm={args.maxm} minm={max(10, args.maxm // 2)} "
                        f"Npass={args.npass} cutoff=1e-10, bonds {args.first_bond}.. (class L, ml=mr={args.maxm})",
            "NT_total": args.nt, "N": 196, "maxm": args.maxm, "Npass": args.npass, "cg_reuse_forward": int(getattr(args, "cg_reuse_forward", 0)),
            "parallelism": f"dp{world} (images sharded, NCCL all-reduce of the bond gradient)" if world > 1 else "dp1",
            "l2": "per-bond inputs (environment cache, >=600 MB/rank at NT=60000) exceed the 126 MB L2; no flush needed"}


# ---------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base, t = cpu_baseline(args, steps=max(1, min(args.steps, 2)))
    line = {"impl": "reference", "metric": "bond-updates/sec", "value": base["value"], "unit": "bond-updates/sec",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 / base["value"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_dict(args, 1),
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "bond-updates/sec", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "images_bonds_per_sec": base["value"] * args.nt}
    print(json.dumps(line))


def run_ours(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from tnml_b200 import capi, fixedl
    feat, labels, W, NTg, first = build_workload(args, rank, world)
    t_setup0 = time.perf_counter()
    h = capi.Handle(local)
    h.set_images(feat, labels, NTg, first)
    h.set_mps(W)
    if world > 1:
        uid = torch.zeros(capi.UNIQUE_ID_BYTES, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid = torch.tensor(list(capi.comm_get_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(uid, 0)
        h.comm_init_rank(world, rank, bytes(uid.cpu().tolist()))
    h.set_option("cg_reuse_forward", args.cg_reuse_forward)
    h.init_envs()
    b0 = args.first_bond
    for bb in range(1, b0):                     # left envs up to the first timed bond
        h.set_bond(bb)
        h.shift_env(bb, capi.FROMLEFT)
    h.synchronize()
    setup_s = time.perf_counter() - t_setup0

    minm = max(10, args.maxm // 2)
    p = capi.BondParams(args.npass, 0.0, 1e-10, 1e-10, args.maxm, minm, 0)
    ext = torch.cuda.ExternalStream(h.stream(), device=torch.device("cuda", local))

    def barrier():
        h.synchronize()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # Bond schedule: class-L bonds with m_l = m_r = maxm only, b = first_bond .. jc-2 going right,
    # then bouncing inside that range like sweepnext does at the chain ends (the turning bond is
    # optimised twice in a row, fixedL.cc:470-476) so that any --steps K can be served.
    lo, hi = b0, 196 // 2 - 2

    def schedule():
        b, ha = lo, 1
        while True:
            yield b, ha
            if ha == 1:
                if b == hi:
                    ha = 2
                else:
                    b += 1
            else:
                if b == lo:
                    ha = 1
                else:
                    b -= 1
    sched = schedule()
    visited = []

    def timed(nsteps, e2e=False):
        """K bond updates; device time by CUDA events on the library's stream."""
        res = []
        h2d = d2h = 0
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(ext):
            ev0.record()
        for k in range(nsteps):
            b, ha = next(sched)
            visited.append(b)
            if e2e:   # the host owns the MPS (like the reference's `W`): sites go H2D, results D2H
                for j in (b, b + 1):
                    Wj = h.get_site(j) if Whost.get(j) is None else Whost[j]
                    h.set_site(j, Wj)
                    h2d += Wj.nbytes
            r = h.bond_update(b, ha, p)
            if e2e:
                for j in (b, b + 1):
                    Whost[j] = h.get_site(j)
                    d2h += Whost[j].nbytes
                d2h += 8 * 40
            res.append(r)
        with torch.cuda.stream(ext):
            ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        if dist is not None:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, res, h2d, d2h

    K, Wm = args.steps, max(args.warmup, 3)
    Whost = {}
    ms, _, _, _ = timed(Wm)                           # warm-up
    uuid = None
    try:
        uuid = "GPU-" + str(torch.cuda.get_device_properties(local).uuid)
    except Exception:
        pass
    sampler = ClockSampler(local, uuid)
    if rank == 0:
        sampler.start()
    h.stats(reset=True)
    ms, res, _, _ = timed(K)
    st = h.stats(reset=True)
    clocks = sampler.finish() if rank == 0 else None
    value = K / (ms / 1000.0)
    timed_bonds = visited[Wm:Wm + K]

    # breakdown pass with per-phase CUDA events (on the library's stream)
    h.set_timing(True)
    ms_t, res_t, _, _ = timed(K)
    stt = h.stats(reset=True)
    h.set_timing(False)

    # end-to-end pass: host-resident MPS, site tensors cross PCIe every step.  The host copies of
    # the sites the pass will touch are fetched before the timed region (they are inputs).
    for j in range(lo, hi + 2):
        Whost[j] = h.get_site(j)
    ms_e, res_e, h2d, d2h = timed(K, e2e=True)
    e2e_value = K / (ms_e / 1000.0)

    value_reuse = None
    if not args.cg_reuse_forward:
        h.set_option("cg_reuse_forward", 1)
        ms_r, _, _, _ = timed(K)
        h.set_option("cg_reuse_forward", 0)
        value_reuse = K / (ms_r / 1000.0)

    if rank == 0:
        peak, peak_src = load_peaks()
        NT = feat.shape[0]
        m = args.maxm
        # CUDA-event breakdown (events recorded on the library's own stream by tnml_set_timing)
        phases = {"proj(krgemm)": stt.ms_proj, "grad(krgram+reduce)": stt.ms_grad, "fat": stt.ms_fat,
                  "svd": stt.ms_svd, "shift": stt.ms_shift, "other": stt.ms_other}
        dom = max(phases, key=phases.get)
        npass = args.npass
        reuse = int(args.cg_reuse_forward)
        n_gemm = ((npass + 2) if reuse else (2 * npass + 1)) * K   # krgemm launches in K bond updates
        n_fwd = (2 * npass + 1) * K          # fat-kernel launches (passes over the fat environment)
        n_bwd = npass * K                    # krgram launches
        # --- dominant data-parallel kernel: krgemm2_kernel<4,3> (FP64 tensor-core MMAs, DMMA.8x8x4)
        # algorithmic flops per launch = 2 * NT * (4*m_l) * m_r  (8 m_l m_r per image)
        gemm_flops_launch = 8.0 * NT * m * m
        gemm_ms_launch = stt.ms_proj / max(1, n_gemm)
        gemm_tf = gemm_flops_launch / (gemm_ms_launch / 1000.0) / 1e12 if gemm_ms_launch > 0 else 0.0
        FP64_TENSOR_PEAK = 37.0   # TF/s, measured on this pool with tools/dmma_bench.cu (DMMA and DFMA share it)
        traffic = None
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_summary.json")))["kernels"]["krgemm2"]
            rd = float(prof["dram__bytes_read.sum"].split()[0]) * (1e6 if "Mbyte" in prof["dram__bytes_read.sum"] else 1e3)
            wr = float(prof["dram__bytes_write.sum"].split()[0]) * (1e6 if "Mbyte" in prof["dram__bytes_write.sum"] else 1e3)
            traffic = (rd + wr) * (NT / 30000.0)       # captured at NT=30000 (ncu cannot replay 62 GB), linear in NT
        except Exception:
            pass
        roof = {"bound": "tensor", "kernel": "krgemm2_kernel<4,3> (Khatri-Rao projection GEMM, FP64 mma.sync m8n8k4, "
                                            "cp.async-staged, persistent)",
                "achieved": gemm_tf, "peak": FP64_TENSOR_PEAK, "unit": "TFLOP/s", "frac": gemm_tf / FP64_TENSOR_PEAK,
                "traffic": traffic,
                "peak_source": "FP64 tensor/FMA pipe measured with tools/dmma_bench.cu on this pool's B200 (37.0 TF/s); "
                               "MEASURED_PEAKS.json holds bf16 and HBM only -- the path computes in f64 (DESIGN.md 3)",
                "launch_avg_ms": gemm_ms_launch, "launches_in_region": n_gemm,
                "algorithmic_flops_per_launch": gemm_flops_launch,
                "phase_ms_per_step": {k: v / K for k, v in phases.items()}, "dominant_phase": dom}
        # --- the HBM-bound kernel of the path: fat_kernel_t (label-carrying environment stream)
        fat_bytes = 8.0 * NT * (m + 10 * m + 1) * n_fwd + 8.0 * NT * m * n_bwd
        fat_gbs = fat_bytes / (stt.ms_fat / 1000.0) / 1e9 if stt.ms_fat > 0 else 0.0
        roof_hbm = {"bound": "hbm", "kernel": "fat_kernel_t", "achieved": fat_gbs, "peak": peak, "unit": "GB/s",
                    "frac": fat_gbs / peak, "peak_source": peak_src, "launch_avg_ms": stt.ms_fat / max(1, n_fwd),
                    "algorithmic_bytes_per_launch": 8.0 * NT * (11 * m + 1)}
        # --- the serial term: truncated SVD of the 2m x 2m bond matrix (replicated on every rank)
        svd_info = {"ms_per_step": stt.ms_svd / K, "sweeps": [int(r.svd_sweeps) for r in res_t][:8],
                    "note": "column sort + 2 Householder QRs + Gram-based block Jacobi (latency bound, ~15 CTAs); "
                            "largest summed share of the step, see profiles/r01_ncu_summary.md"}
        line = {"metric": "bond-updates/sec", "value": value, "unit": "bond-updates/sec", "n_gpus": world,
                "steps": K, "warmup": Wm, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_dict(args, world),
                "images_bonds_per_sec": value * NTg,
                "e2e": {"value": e2e_value, "unit": "bond-updates/sec", "h2d_bytes_per_step": h2d / K,
                        "d2h_bytes_per_step": d2h / K,
                        "note": "host-resident MPS: W(b),W(b+1) H2D before and D2H after every bond update "
                                "through the C-ABI; images/environments are resident state (TrainStates)"},
                "gpu_launches": int(st.launches), "clocks": clocks, "roofline": roof, "roofline_hbm": roof_hbm,
                "svd": svd_info,
                "value_cg_reuse_forward": value_reuse,
                "setup_s": setup_s, "newm": [int(r.newm) for r in res][:8], "newm_min": min(int(r.newm) for r in res),
                "timed_bonds": [int(timed_bonds[0]), int(timed_bonds[-1])],
                "cost_per_image": [r.cost / NTg for r in res][:4]}
        if not args.no_cpu_baseline:
            base, _ = cpu_baseline(args, steps=1)
            line["cpu_baseline"] = base
        print(json.dumps(line))
    h.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
