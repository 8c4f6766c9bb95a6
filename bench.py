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
    ap.add_argument("--config", type=int, default=3, choices=[3, 5],
                    help="BASELINE config: 3 (= 4 with --gpus N; default) full-MNIST-shaped 60000 x 196 sites, maxm=120; "
                         "5: synthetic 1e6 images, maxm=300, window of bonds on an 8-site chain")
    ap.add_argument("--cpu-sample", type=int, default=1024, help="images in the CPU-baseline sample (literal dense t.v)")
    ap.add_argument("--cpu-sample-structured", type=int, default=4000,
                    help="images in the structured (Khatri-Rao form, BLAS) CPU-baseline sample")
    ap.add_argument("--sweep-avg", type=int, default=1, help="1: also time one full sweep (config 3)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cg-reuse-forward", type=int, default=0,
                    help="1: linear update of the forward outputs between CG passes (tnml_set_option); "
                         "0 (default): literal recompute like fixedL.cc:412-421")
    a = ap.parse_args()
    if a.config == 5:
        if a.nt == 60000:
            a.nt = 1000000
        if a.maxm == 120:
            a.maxm = 300
        a.first_bond = 2
    return a


# ---------------------------------------------------------------------------
def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_tensor_peaks():
    """Tensor-pipe peaks measured on THIS box right before the timed run (rank 0): the FP64 DMMA pipe
    (tools/dmma_bench) and the tcgen05 int8 pipe at the clock the power cap allows (tools/umma_bench).
    MEASURED_PEAKS.json (driver-written) holds only HBM and bf16; both are recorded next to these."""
    out = {}
    pj = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pj):
        j = json.load(open(pj))
        out["bf16_tflops_file"] = float(j.get("bf16_tflops", 0.0))
        out["bf16_tflops_sustained_file"] = float(j.get("bf16_tflops_sustained", 0.0))
        out["hbm_gbs_file"] = float(j.get("hbm_gbs", 0.0))
    import re
    exe = os.path.join(ROOT, "tools", "dmma_bench")
    if os.path.exists(exe):
        try:
            o = subprocess.run([exe], capture_output=True, text=True, timeout=60).stdout
            mt = re.search(r"dmma only.*?DMMA ([0-9.]+) TF/s", o)
            if mt:
                out["dmma_tflops"] = float(mt.group(1))
                out["dmma_src"] = "tools/dmma_bench in this run (mma.sync.m8n8k4.f64, all SMs)"
        except Exception:
            pass
    exe = os.path.join(ROOT, "tools", "umma_bench")
    if os.path.exists(exe):
        try:
            o = subprocess.run([exe], capture_output=True, text=True, timeout=60).stdout
            mt = re.search(r"INT8_PEAK_TOPS ([0-9.]+).*?SM clock ([0-9.]+) MHz", o)
            if mt:
                out["int8_tops"] = float(mt.group(1))
                out["int8_sm_mhz"] = float(mt.group(2))
                out["int8_src"] = ("tools/umma_bench in this run: tcgen05.mma kind::i8 M=128 N=256 back to back on all SMs, "
                                   "wall clock (power-capped SM clock %.0f MHz)" % float(mt.group(2)))
        except Exception:
            pass
    return out


def load_traffic(NT):
    """dram bytes per launch of the dominant kernel from the committed ncu capture of this round
    (profiles/r02_ncu_summary.json, written by tools/make_profiles.py with the commit it was taken at);
    captured at a smaller NT (ncu replays every kernel ~40 times) and scaled linearly in NT."""
    try:
        j = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_summary.json")))
        k = j["kernels"]["oz_gemm"]
        return float(k["dram_bytes"]) * (NT / float(k["NT"])), f"ncu --set full at NT={k['NT']} (commit {j.get('commit', '?')}), scaled"
    except Exception:
        return None, "no ncu capture of this round committed yet"


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


def chain_geometry(args):
    """(N sites, first / last saturated class-L bond, bonds of one timed excursion builder)."""
    if args.config == 5:
        return 8
    return 196


def build_workload(args, rank, world):
    """Synthetic inputs for this rank's shard (ParallelDo bounds of the global image list)."""
    from tnml_b200 import data, fixedl
    NTg = args.nt
    b0, b1 = fixedl.bounds(world, NTg)[rank]
    if args.config == 5:
        pix, labels = data.synthetic_pixels(b1 - b0, 8, seed=20260925, first=b0)
        feat = data.phi(pix)                  # same feature map as config 3 (pixels / 255, phi divides by 255 again)
        W = data.window_mps(8, 2, args.maxm, seed=5)
        return feat, labels, W, NTg, b0
    pix, labels = data.synthetic_digits(b1 - b0, 14, seed=20260925, first=b0)
    feat = data.phi(pix)                      # [NT, 196, 2] = TState::data
    W = data.random_mps(196, 2, args.maxm, seed=3)
    return feat, labels, W, NTg, b0


def excursion(args, K):
    """Bond schedule of one timed pass of K bond updates that returns to its starting state, so that
    the pass can be repeated (main / breakdown / end-to-end / cg_reuse_forward passes).
    config 3: class-L bonds with m_l = m_r = maxm: ceil(K/2) bonds rightwards from first_bond (ha=1), then
    floor(K/2) leftwards back (ha=2; the turning bond is optimised twice in a row like sweepnext does at
    the chain end, fixedL.cc:470-476): rightward bonds advance a thin environment, leftward ones the
    label-carrying one (kappa = 10), so both halves of a sweep are in the figure.
    config 5: the window 2..6 of the 8-site chain (class L, C, C, R, R) rightwards then leftwards."""
    if args.config == 5:
        cyc = [(b, 1) for b in range(2, 7)] + [(b, 2) for b in range(6, 1, -1)]
        return [cyc[k % len(cyc)] for k in range(K)]
    lo = args.first_bond
    nr = (K + 1) // 2
    right = [(lo + k, 1) for k in range(nr)]
    left = [(lo + nr - 1 - k, 2) for k in range(K - nr)]
    return right + left


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


def cpu_baseline_structured(args):
    """All host cores for the BLAS calls even when a launcher pinned OMP_NUM_THREADS=1 (torchrun does)."""
    try:
        from threadpoolctl import threadpool_limits
        with threadpool_limits(limits=os.cpu_count()):
            return _cpu_baseline_structured(args)
    except ImportError:
        return _cpu_baseline_structured(args)


def _cpu_baseline_structured(args):
    """BASELINE.md variant (ii): the best-case CPU formulation -- the same Khatri-Rao-structured
    contractions the GPU path uses, as BLAS GEMMs (numpy/OpenBLAS, all host cores) in the oracle -- one
    whole bond update (cgrad + svd + quadcost + shiftE) on a sample of the images, scaled linearly in NT
    (the SVD, which does not scale with NT, is timed separately and added unscaled)."""
    from oracle import fixedl_oracle as O           # checker / baseline only
    from tnml_b200 import data
    ns = min(args.cpu_sample_structured, args.nt)
    if args.config == 5:
        pix, labels = data.synthetic_pixels(ns, 8, seed=20260925)
        feat = O.features(pix)
        W = data.window_mps(8, 2, args.maxm, seed=5)
        b = 2
    else:
        # a class-L bond with ml = mr = maxm costs the same on a 24-site chain as on the 196-site one;
        # the short chain keeps the (untimed) environment set-up of the sample small
        pix, labels = data.synthetic_digits(ns, 14, seed=20260925)
        feat = O.features(pix[:, 86:86 + 24])
        W = data.random_mps(24, 2, args.maxm, seed=3)
        b = 10
    chunk = 1000
    tss = [O.TrainStates(feat[a:a + chunk], labels[a:a + chunk].astype(np.int64)) for a in range(0, ns, chunk)]
    for t in tss:
        t.init(W)
        for bb in range(1, b):
            t.set_bond(bb)
            t.shiftE(W, bb, "Fromleft")
        t.set_bond(b)
    B = O.form_bond(W[b], W[b + 1])
    jc = tss[0].jc

    def grad(X):
        G, C = np.zeros_like(X), 0.0
        for t in tss:
            g, c = O._grad(X, t, 0.0, False)
            G, C = G + g, C + c
        return G, C
    t0 = time.perf_counter()
    r, _ = grad(B)
    p_ = r.copy()
    for ps in range(1, args.npass + 1):
        pAp = sum(float(np.sum(O.project(p_, t) ** 2)) for t in tss)
        B = B + (float(np.sum(r * r)) / pAp) * p_
        if ps == args.npass:
            break
        nr, _ = grad(B)
        beta = float(np.sum(nr * nr)) / float(np.sum(r * r))
        r = nr
        p_ = r + beta * p_
    t_cg = time.perf_counter() - t0
    t0 = time.perf_counter()
    Wb, Wb1, m, te = O.svd_split(B, b, 1, jc, args.maxm, max(10, args.maxm // 2), 1e-10)
    t_svd = time.perf_counter() - t0
    t0 = time.perf_counter()
    newB = O.form_bond(Wb, Wb1)
    for t in tss:
        O.quadcost(newB, t)
    Wn = list(W)
    Wn[b], Wn[b + 1] = Wb, Wb1
    for t in tss:
        t.shiftE(Wn, b, "Fromleft")
    t_rest = time.perf_counter() - t0
    t_full = (t_cg + t_rest) * args.nt / ns + t_svd
    return {"value": 1.0 / t_full, "unit": "bond-updates/sec", "cores": os.cpu_count(), "kind": "port",
            "sample": f"{ns} images, one whole class-L bond update at ml=mr={args.maxm} in the structured (Khatri-Rao, BLAS GEMM) "
                      f"form of the numpy oracle: cgrad {t_cg:.2f} s + quadcost/shiftE {t_rest:.2f} s scaled linearly to "
                      f"NT={args.nt}, + LAPACK svd {t_svd:.3f} s unscaled; numpy/OpenBLAS threads = host cores"}


def config_dict(args, world):
    if args.config == 5:
        wl = ("BASELINE config 5: synthetic hash-RNG pixels, d=2 NL=10, "
              f"NT={args.nt} images total, maxm=minm={args.maxm} Npass={args.npass}, window of bonds 2..6 (class L, C, C, R, R; "
              f"rightwards then leftwards) of an 8-site chain with link dimension {args.maxm} on every bond")
        N = 8
    else:
        wl = ("BASELINE config 3/4: synthetic 14x14 MNIST-shaped, N=196 sites d=2 NL=10, "
              f"NT={args.nt} images total, maxm={args.maxm} minm={max(10, args.maxm // 2)} "
              f"Npass={args.npass} cutoff=1e-10, class-L bonds with ml=mr={args.maxm}: half of the steps rightwards from bond "
              f"{args.first_bond}, half leftwards back")
        N = 196
    return {"workload": wl,
            "NT_total": args.nt, "N": N, "maxm": args.maxm, "Npass": args.npass, "cg_reuse_forward": int(getattr(args, "cg_reuse_forward", 0)),
            "parallelism": f"dp{world} (images sharded, NCCL all-reduce of the bond gradient)" if world > 1 else "dp1",
            "l2": "per-bond inputs (environment cache, >=600 MB/rank at NT=60000) exceed the 126 MB L2; no flush needed"}


# ---------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.config == 5:      # dense t.v at m=300 is 29 MB per image: only the structured form is runnable
        base = cpu_baseline_structured(args)
    else:
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

    minm = args.maxm if args.config == 5 else max(10, args.maxm // 2)
    p = capi.BondParams(args.npass, 0.0, 1e-10, 1e-10, args.maxm, minm, 0)
    ext = torch.cuda.ExternalStream(h.stream(), device=torch.device("cuda", local))

    def barrier():
        h.synchronize()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(nsteps, e2e=False):
        """One excursion of nsteps bond updates (it returns to the starting state); device time by CUDA
        events on the library's stream, also for the rightward and the leftward half."""
        sched = excursion(args, nsteps)
        res = []
        h2d = d2h = 0
        barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        turn = next((k for k, (_, ha) in enumerate(sched) if ha == 2), len(sched))
        with torch.cuda.stream(ext):
            ev[0].record()
        for k, (b, ha) in enumerate(sched):
            if k == turn:
                with torch.cuda.stream(ext):
                    ev[1].record()
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
            if turn >= len(sched):
                ev[1].record()
            ev[2].record()
        barrier()
        ms = [ev[0].elapsed_time(ev[2]), ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])]
        if dist is not None:
            t = torch.tensor(ms, device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = [float(x) for x in t.tolist()]
        return ms, res, h2d, d2h, sched

    K, Wm = args.steps, max(args.warmup, 3)
    if args.config == 5:          # whole window sweeps only (the excursion must return to its start)
        K = max(10, ((K + 9) // 10) * 10)
        Wm = 10
    Whost = {}
    timed(Wm)                                         # warm-up
    uuid = None
    try:
        uuid = "GPU-" + str(torch.cuda.get_device_properties(local).uuid)
    except Exception:
        pass
    sampler = ClockSampler(local, uuid)
    if rank == 0:
        sampler.start()
    h.stats(reset=True)
    ms3, res, _, _, sched = timed(K)
    st = h.stats(reset=True)
    clocks = sampler.finish() if rank == 0 else None
    ms = ms3[0]
    value = K / (ms / 1000.0)
    n_right = sum(1 for (_, ha) in sched if ha == 1)
    n_left = K - n_right

    # breakdown pass with per-phase CUDA events (on the library's stream)
    h.set_timing(True)
    _, res_t, _, _, _ = timed(K)
    stt = h.stats(reset=True)
    h.set_timing(False)

    # end-to-end pass: host-resident MPS, site tensors cross PCIe every step.  The host copies of
    # the sites the pass will touch are fetched before the timed region (they are inputs).
    for j in sorted({x for (b, _) in sched for x in (b, b + 1)}):
        Whost[j] = h.get_site(j)
    mse, res_e, h2d, d2h, _ = timed(K, e2e=True)
    e2e_value = K / (mse[0] / 1000.0)

    value_reuse = None
    if not args.cg_reuse_forward:
        h.set_option("cg_reuse_forward", 1)
        msr, _, _, _, _ = timed(K)
        h.set_option("cg_reuse_forward", 0)
        value_reuse = K / (msr[0] / 1000.0)

    # one FULL sweep (2(N-1) bond updates incl. the cheap edge bonds and the expensive class-C bonds)
    # from the initial MPS: the user-visible sweep-average rate next to the saturated-bond rate above
    sweep_avg = None
    if args.sweep_avg and args.config == 3:
        h.set_mps(W)
        h.init_envs()
        times = []
        last = None
        for sw in range(2):          # the first sweep sizes every buffer (class-C work arrays, planes); the second is reported
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(ext):
                e0.record()
            nb = 0
            for b, ha in fixedl.sweepnext(196):
                last = h.bond_update(b, ha, p)
                nb += 1
            with torch.cuda.stream(ext):
                e1.record()
            barrier()
            mss = e0.elapsed_time(e1)
            if dist is not None:
                t = torch.tensor([mss], device="cuda", dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                mss = float(t.item())
            times.append(mss / 1000.0)
        sweep_avg = {"value": nb / times[1], "unit": "bond-updates/sec", "bond_updates": nb, "seconds": times[1],
                     "first_sweep_seconds": times[0],
                     "note": "second of two full sweeps b=1..195..1 from the initial random MPS (link dims min(2^j, 2^(N-j), maxm)): "
                             "includes the small edge bonds and the four class-C bonds (label index on the bond tensor); the first "
                             "sweep also sizes every work buffer",
                     "final_cost_per_image": last.cost / NTg}

    if rank == 0:
        peak, peak_src = load_peaks()
        peaks = measured_tensor_peaks()
        NT = feat.shape[0]
        m = args.maxm
        phases = {"proj(slice+oz_gemm)": stt.ms_proj, "grad(krgram+reduce)": stt.ms_grad, "fat": stt.ms_fat,
                  "svd": stt.ms_svd, "shift": stt.ms_shift, "other": stt.ms_other}
        dom = max(phases, key=phases.get)
        npass = args.npass
        reuse = int(args.cg_reuse_forward)
        nclsL = sum(1 for (b, _) in sched if not (args.config == 5 and b in (3, 4)))   # class L/R bond updates
        n_gemm = ((npass + 2) if reuse else (2 * npass + 1)) * K   # projection launches in K bond updates
        n_fwd = (2 * npass + 1) * K          # fat-kernel launches (passes over the fat environment)
        n_bwd = npass * K                    # krgram launches
        # --- dominant data-parallel kernel: oz_gemm_kernel<8,4> (tcgen05.mma kind::i8, TMA, TMEM)
        # algorithmic float64 work per launch = 2 * NT * (4 m_l) * m_r flop (8 m_l m_r per image);
        # executed tensor work = 36 exact int8 slice products on K padded to 128: 2 * NTpad * 128 * 4 m_r * 36 ops
        gemm_flops_launch = 8.0 * NT * m * m
        ntp = ((NT + 127) // 128) * 128
        gemm_int8_ops_launch = 2.0 * ntp * 128.0 * (4.0 * ((m + 7) // 8) * 8) * 36.0
        gemm_ms_launch = stt.ms_proj / max(1, n_gemm)     # includes the plane cutting (once per bond + once per launch)
        tops = gemm_int8_ops_launch / (gemm_ms_launch / 1000.0) / 1e12 if gemm_ms_launch > 0 else 0.0
        # fallback when tools/umma_bench is not built: the nominal dense int8 rate (4.5 POP/s), never below 2 x the bf16 file figure
        i8_peak = peaks.get("int8_tops") or max(4500.0, 2.0 * peaks.get("bf16_tflops_file", 0.0))
        traffic, traffic_src = load_traffic(NT)
        tc_path = (m <= 128 and NT >= 1024)
        if tc_path:
            roof = {"bound": "tensor", "kernel": "oz_gemm_kernel<8,4> (Khatri-Rao projection GEMM as 36 exact int8 slice products: "
                                                "tcgen05.mma kind::i8 + TMA + TMEM, float64-class result)",
                    "achieved": tops, "peak": i8_peak, "unit": "TOP/s (int8 tensor ops executed; TFLOP/s-equivalent below)",
                    "frac": tops / i8_peak if i8_peak else None, "traffic": traffic, "traffic_source": traffic_src,
                    "peak_source": peaks.get("int8_src", "fallback: nominal dense int8 4500 TOP/s (tools/umma_bench not built; "
                                                      "MEASURED_PEAKS.json has no int8 entry)"),
                    "fp64_equiv_tflops": gemm_flops_launch / (gemm_ms_launch / 1000.0) / 1e12 if gemm_ms_launch > 0 else 0.0,
                    "fp64_dmma_pipe_tflops": peaks.get("dmma_tflops"),
                    "launch_avg_ms": gemm_ms_launch, "launches_in_region": n_gemm,
                    "algorithmic_flops_per_launch": gemm_flops_launch, "int8_ops_per_launch": gemm_int8_ops_launch,
                    "phase_ms_per_step": {k: v / K for k, v in phases.items()}, "dominant_phase": dom}
        else:
            gemm_tf = gemm_flops_launch / (gemm_ms_launch / 1000.0) / 1e12 if gemm_ms_launch > 0 else 0.0
            dpk = peaks.get("dmma_tflops") or 37.0
            roof = {"bound": "tensor", "kernel": "krgemm_kernel / krgemm2_kernel (FP64 mma.sync m8n8k4; link dimension > 128: "
                                                "the tcgen05 kernel keeps K <= 128 resident)",
                    "achieved": gemm_tf, "peak": dpk, "unit": "TFLOP/s", "frac": gemm_tf / dpk, "traffic": None,
                    "peak_source": "FP64 DMMA pipe measured in this run (tools/dmma_bench)" if peaks.get("dmma_tflops") else
                                   "37.0 TF/s (tools/dmma_bench, round 1)",
                    "launch_avg_ms": gemm_ms_launch, "launches_in_region": n_gemm,
                    "algorithmic_flops_per_launch": gemm_flops_launch,
                    "phase_ms_per_step": {k: v / K for k, v in phases.items()}, "dominant_phase": dom}
        # --- the HBM-bound kernel of the path: fat_kernel_t (label-carrying environment stream)
        fat_bytes = 8.0 * NT * (m + 10 * m + 1) * n_fwd + 8.0 * NT * m * n_bwd
        fat_gbs = fat_bytes / (stt.ms_fat / 1000.0) / 1e9 if stt.ms_fat > 0 else 0.0
        roof_hbm = {"bound": "hbm", "kernel": "fat_kernel_t", "achieved": fat_gbs, "peak": peak, "unit": "GB/s",
                    "frac": fat_gbs / peak, "peak_source": peak_src, "launch_avg_ms": stt.ms_fat / max(1, n_fwd),
                    "algorithmic_bytes_per_launch": 8.0 * NT * (11 * m + 1),
                    "note": "class L/R bonds only" if args.config == 3 else "window incl. class-C bonds: indicative only"}
        svd_info = {"ms_per_step": stt.ms_svd / K, "sweeps": [int(r.svd_sweeps) for r in res_t][:10],
                    "note": "column sort + 2 Householder QRs + cluster-resident block Jacobi (latency bound, 16 CTAs), replicated on every rank"}
        line = {"metric": "bond-updates/sec", "value": value, "unit": "bond-updates/sec", "n_gpus": world,
                "steps": K, "warmup": Wm, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_dict(args, world),
                "images_bonds_per_sec": value * NTg,
                "value_rightward": n_right / (ms3[1] / 1000.0) if ms3[1] > 0 else None,
                "value_leftward": n_left / (ms3[2] / 1000.0) if n_left and ms3[2] > 0 else None,
                "value_sweep_avg": sweep_avg,
                "e2e": {"value": e2e_value, "unit": "bond-updates/sec", "h2d_bytes_per_step": h2d / K,
                        "d2h_bytes_per_step": d2h / K,
                        "note": "host-resident MPS: W(b),W(b+1) H2D before and D2H after every bond update "
                                "through the C-ABI; images/environments are resident state (TrainStates)"},
                "gpu_launches": int(st.launches), "clocks": clocks, "roofline": roof, "roofline_hbm": roof_hbm,
                "svd": svd_info, "measured_peaks": peaks,
                "value_cg_reuse_forward": value_reuse,
                "setup_s": setup_s, "newm": [int(r.newm) for r in res][:10], "newm_min": min(int(r.newm) for r in res),
                "timed_bonds": [[int(b), int(ha)] for (b, ha) in sched][:24],
                "cost_per_image": [r.cost / NTg for r in res][:4]}
        if not args.no_cpu_baseline:
            if args.config == 3:
                base, _ = cpu_baseline(args, steps=1)
                line["cpu_baseline"] = base
            try:
                line["cpu_baseline_structured"] = cpu_baseline_structured(args)
                if args.config == 5:
                    line["cpu_baseline"] = line["cpu_baseline_structured"]
            except Exception as e:   # a baseline problem must not lose the GPU measurement
                line["cpu_baseline_structured"] = {"error": repr(e)}
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
