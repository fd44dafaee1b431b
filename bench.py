#!/usr/bin/env python
"""bench.py — LF Mpix/s of both LFBM5D steps on a synthetic 17x17x1024^2 RGB light field (BASELINE.json configs[2]).

  python bench.py --gpus N --steps K --warmup W           # this framework, one rank per GPU
  python bench.py --impl reference --gpus N ...            # the reference's own CPU code on the host cores

A "step" is one full denoising of the light field: hard-threshold step then Wiener step. `value` is measured with
the light field already resident in HBM (device pointers into the C ABI); `e2e` goes through the host-buffer C ABI
calls a drop-in user makes (pinned host arrays, H2D/D2H inside the timed region). Rank 0 prints ONE JSON line.
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
sys.path.insert(0, os.path.join(ROOT, "tests"))

# BASELINE.json configs[2] / SURVEY.md 8(d) "Config 3": README.md:61 parameters
CFG = dict(aw=17, ah=17, H=1024, W=1024, C=3, sigma=10.0, lam=2.7, an=1,
           s1=dict(N=8, nSim=18, nDisp=6, k=16, p=4, tau2="id"), s2=dict(N=16, nSim=18, nDisp=6, k=8, p=4, tau2="dct"))
WORKLOAD = "configs[2]: Stanford-shaped synthetic LF 17x17 SAIs 1024x1024 RGB, sigma 10, 8 18 6 16 4 id sadct haar / 16 18 6 8 4 dct sadct haar, opp"
N_PASSES = 64            # window passes per step at 17x17 with an = 1


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, False, []

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.check_output(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                               "--format=csv,noheader,nounits"], timeout=5).decode().strip()
                self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]) if self.rows[0][1].replace(".", "").isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


def params(L, which):
    s = CFG[which]
    tau2 = {"id": L.ID, "dct": L.DCT, "bior": L.BIOR}[s["tau2"]]
    return L.make_params(CFG["sigma"], CFG["lam"] if which == "s1" else 0.0, CFG["aw"], CFG["ah"], CFG["an"], CFG["W"], CFG["H"], CFG["C"],
                         s["N"], s["nSim"], s["nDisp"], s["k"], s["p"], tau2, L.SADCT, L.HAAR)


def grid_counts(which):
    s = CFG[which]
    n = s["nSim"] + s["nDisp"]

    def cnt(dim):
        m = dim + 2 * n - s["k"] + 1
        idx = list(range(n, m - n, s["p"]))
        if idx[-1] < m - n - 1:
            idx.append(m - n - 1)
        return len(idx)
    return cnt(CFG["H"]) * cnt(CFG["W"])


def algorithmic_bytes_groups(which):
    """unfused two-kernel model, 2*G*4 B per window pass, G = R*N*A*C*k^2 coefficients."""
    s = CFG[which]
    G = grid_counts(which) * s["N"] * 9 * CFG["C"] * s["k"] ** 2
    return 2.0 * G * 4.0


def algorithmic_flops_bm(which):
    """direct-SSD count, 3 flop per pixel pair: self R*(2nSim+1)^2*k^2*3 + stereo (A-1)*min(P,R*N)*(2nDisp+1)^2*k^2*3."""
    s = CFG[which]
    n = s["nSim"] + s["nDisp"]
    R = grid_counts(which)
    P = (CFG["H"] + 2 * n - s["k"] + 1 - 2 * s["nDisp"]) * (CFG["W"] + 2 * n - s["k"] + 1 - 2 * s["nDisp"])
    return R * (2 * s["nSim"] + 1) ** 2 * s["k"] ** 2 * 3.0 + 8 * min(P, R * s["N"]) * (2 * s["nDisp"] + 1) ** 2 * s["k"] ** 2 * 3.0


def run_ours(args):
    import torch
    import lfbm5d_b200 as L
    import lfdata

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    eng = L.LFBM5D(local)
    stream = torch.cuda.ExternalStream(eng.stream(), device=dev)
    asize, H, W, C = CFG["aw"] * CFG["ah"], CFG["H"], CFG["W"], CFG["C"]
    lf_pix = asize * H * W
    # synthetic clean LF: procedural texture cropped at 1 px/view disparity; Gaussian noise sigma 10 (seeded per rank)
    pad = CFG["aw"]
    base = torch.from_numpy(lfdata.base_image(H + 2 * pad, W + 2 * pad, C, seed=12345 + rank)).to(dev)
    clean = torch.empty((asize, C, H, W), device=dev)
    for s_ in range(CFG["ah"]):
        for t_ in range(CFG["aw"]):
            oy, ox = pad + (s_ - CFG["ah"] // 2), pad + (t_ - CFG["aw"] // 2)
            clean[s_ * CFG["aw"] + t_] = base[:, oy:oy + H, ox:ox + W]
    gen = torch.Generator(device=dev)
    gen.manual_seed(20171016 + rank)
    noisy0 = clean + CFG["sigma"] * torch.randn(clean.shape, device=dev, generator=gen)
    work, basic, out = torch.empty_like(noisy0), torch.empty_like(noisy0), torch.empty_like(noisy0)
    mask = np.ones(asize, np.uint32)
    p1, p2 = params(L, "s1"), params(L, "s2")
    if args.passes:
        eng.set_max_passes(args.passes)
    torch.cuda.synchronize()

    def one_step():
        with torch.cuda.stream(stream):
            work.copy_(noisy0, non_blocking=True)
        eng.step1_device(p1, work.data_ptr(), mask, basic.data_ptr())
        eng.step2_device(p2, work.data_ptr(), basic.data_ptr(), mask, out.data_ptr())

    for _ in range(args.warmup):
        one_step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    eng.reset_stats()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        one_step()
    e1.record(stream)
    e1.synchronize()
    barrier()
    ms = e0.elapsed_time(e1)
    st = eng.stats()
    launches = int(st.kernel_launches)
    sampler.stop_flag = True
    tmax = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms = float(tmax.item())
    ms_per_step = ms / args.steps
    frac_passes = 1.0 if not args.passes else args.passes / float(N_PASSES)
    value = world * lf_pix * frac_passes / (ms_per_step * 1e-3) / 1e6

    # quality of the timed output (informational): PSNR against the clean LF
    mse = float(((out - clean) ** 2).mean().item())
    psnr_out = 20.0 * np.log10(255.0 / np.sqrt(mse)) if not args.passes else None
    psnr_in = 20.0 * np.log10(255.0 / np.sqrt(float(((noisy0 - clean) ** 2).mean().item())))

    # ---- e2e: host-buffer C ABI with pinned memory, copies inside the timed region ----
    e2e = None
    e2e_ok = 0
    if not args.no_e2e:
        try:        # 3 x 3.6 GB of pinned host memory per rank; every rank has to get it, or all skip the measurement
            h_noisy = torch.empty(noisy0.shape, pin_memory=True)
            h_basic = torch.empty(noisy0.shape, pin_memory=True)
            h_out = torch.empty(noisy0.shape, pin_memory=True)
            h_noisy0 = noisy0.cpu()
            e2e_ok = 1
        except RuntimeError:
            e2e_ok = 0
        if world > 1:
            f = torch.tensor([e2e_ok], device=dev)
            dist.all_reduce(f, op=dist.ReduceOp.MIN)
            e2e_ok = int(f.item())
        if not e2e_ok:
            e2e = {"value": None, "unit": "LF Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "error": "pinned host allocation failed"}
    if e2e_ok:
        del clean
        torch.cuda.empty_cache()

        def ptrs(t):
            import ctypes as Cc
            each = t[0].numel()
            return (Cc.POINTER(Cc.c_float) * asize)(*[Cc.cast(t.data_ptr() + 4 * each * i, Cc.POINTER(Cc.c_float)) for i in range(asize)])

        import ctypes as Cc
        m = mask.ctypes.data_as(Cc.POINTER(Cc.c_uint))

        def e2e_step():
            h_noisy.copy_(h_noisy0)
            if eng.lib.lfbm5d_step1(eng.ctx, Cc.byref(p1), ptrs(h_noisy), m, ptrs(h_basic)) != 0:
                raise RuntimeError(eng.error())
            if eng.lib.lfbm5d_step2(eng.ctx, Cc.byref(p2), ptrs(h_noisy), ptrs(h_basic), m, ptrs(h_out)) != 0:
                raise RuntimeError(eng.error())
            return float(h_out[0, 0, 0, 0])      # the result is read on the host

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        torch.cuda.synchronize()
        t_e2e = torch.tensor([time.perf_counter() - t0], device=dev)
        if world > 1:
            dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
        nbytes = noisy0.numel() * 4
        e2e = {"value": world * lf_pix * frac_passes / (float(t_e2e.item()) / args.steps) / 1e6, "unit": "LF Mpix/s",
               "h2d_bytes_per_step": 3 * nbytes, "d2h_bytes_per_step": 5 * nbytes, "timer": "host wall clock around the C ABI calls, max over ranks"}

    # ---- strong scaling (N > 1): ONE light field on all ranks, windows that share no SAI on different GPUs, accumulators
    # broadcast over NCCL per plan level (lfbm5d_b200/dist.py); bit-identical to the single-GPU result. Reported beside the
    # weak-scaling headline, not instead of it.
    strong = None
    if world > 1 and not args.passes and not args.no_strong:
        from lfbm5d_b200 import dist as D
        t_strong = []
        for it in range(2):                                   # one warm-up (buffers, NCCL), one timed
            work.copy_(noisy0)
            barrier()
            t0 = time.perf_counter()
            D.run_step_windows(eng, 1, p1, work.data_ptr(), 0, mask, basic.data_ptr(), dist, dev)
            plan = D.run_step_windows(eng, 2, p2, work.data_ptr(), basic.data_ptr(), mask, out.data_ptr(), dist, dev)
            barrier()
            t_strong.append(time.perf_counter() - t0)
        ts = torch.tensor([t_strong[-1]], device=dev)
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        strong = {"value": lf_pix / float(ts.item()) / 1e6, "unit": "LF Mpix/s", "seconds": float(ts.item()),
                  "speedup_vs_one_gpu": (ms_per_step * 1e-3) / float(ts.item()), "plan_levels": int(plan[:, 4].max()) + 1,
                  "windows": int(len(plan)), "note": "one light field, window-level parallelism, results identical to one GPU"}

    # ---- per-kernel timing for the roofline (separate short run with per-phase CUDA events on the library's stream) ----
    roof, roof_other, roof_bm, phases = None, None, None, None
    if rank == 0 and args.profile_passes > 0:
        eng.enable_timing(True)
        eng.set_max_passes(args.profile_passes)
        eng.reset_stats()
        eng.step1_device(p1, work.data_ptr(), mask, basic.data_ptr())
        s1 = eng.stats()
        eng.reset_stats()
        eng.step2_device(p2, work.data_ptr(), basic.data_ptr(), mask, out.data_ptr())
        s2 = eng.stats()
        eng.enable_timing(False)
        eng.set_max_passes(args.passes or 0)
        hbm, how = peaks()
        n1, n2 = max(1, s1.window_passes), max(1, s2.window_passes)
        g_ms = (s1.ms_groups / n1, s2.ms_groups / n2)
        a_ms = (s1.ms_aggregate / n1, s2.ms_aggregate / n2)
        sat_ms = (s1.ms_sat / n1, s2.ms_sat / n2)
        # SURVEY 8(d) two-kernel model: the transform kernel writes G*4 B of filtered coefficients (and reads the 9 padded
        # SAIs once), the aggregation kernel reads G*4 B and writes num/den once
        half = (algorithmic_bytes_groups("s1") / 2.0, algorithmic_bytes_groups("s2") / 2.0)
        wb = (CFG["W"] + 48) * (CFG["H"] + 48) * 9 * CFG["C"] * 4.0
        bytes_t = (half[0] + wb, half[1] + 2 * wb)
        bytes_a = (half[0] + 2 * wb, half[1] + 2 * wb)
        traffic = ncu_traffic()

        def entry(kernel, ncu_name, nbytes, ms, model):
            t = traffic.get(ncu_name)
            ach = nbytes / (ms * 1e-3) / 1e9
            return {"kernel": kernel, "bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm,
                    "traffic": None if t is None else t["dram_gb_per_launch"] * 1e9, "traffic_source": None if t is None else traffic["_file"],
                    "algorithmic_bytes": nbytes, "ms_per_launch": ms, "peak_source": how, "model": model}

        m_t = "algorithmic bytes per launch = G*4 B written + padded window read once (SURVEY 8(d)), G=R*N*A*C*k^2"
        m_a = "algorithmic bytes per launch = G*4 B read + num/den read and written once (SURVEY 8(d))"
        kernels = [entry("k_groups_id16<3> (step 1)", "k_groups_id16<3>", bytes_t[0], g_ms[0], m_t),
                   entry("k_groups_w8<3> (step 2)", "k_groups_w8<3>", bytes_t[1], g_ms[1], m_t),
                   entry("k_aggregate<16, 3> (step 1)", "k_aggregate<16, 3>", bytes_a[0], a_ms[0], m_a),
                   entry("k_aggregate<8, 3> (step 2)", "k_aggregate<8, 3>", bytes_a[1], a_ms[1], m_a)]
        kernels[1]["note"] = "issue-limited, not bandwidth-limited: ~100 flop per coefficient (2-D + angular DCT, Haar, Wiener, inverses)"
        kernels.sort(key=lambda e: -e["ms_per_launch"])
        roof, roof_other = kernels[0], kernels[1:]
        fl = (algorithmic_flops_bm("s1"), algorithmic_flops_bm("s2"))
        ach_bm = (fl[0] + fl[1]) / ((sat_ms[0] + sat_ms[1]) * 1e-3) / 1e12
        roof_bm = {"kernel": "k_sat_planes", "bound": "fp32", "achieved": ach_bm, "peak": 74.4, "unit": "TFLOP/s", "frac": ach_bm / 74.4,
                   "peak_source": "nominal 148 SM x 128 lanes x 2 x 1.965 GHz (no measured FP32 peak)",
                   "model": "direct-SSD-equivalent flops (SURVEY 8(d)); the kernel itself runs the reference's summed-area recurrence",
                   "ms_per_pass": {"step1": sat_ms[0], "step2": sat_ms[1]}}
        phases = {"step1_ms_per_pass": {"block_matching": s1.ms_block_matching / n1, "groups": g_ms[0], "aggregate": a_ms[0]},
                  "step2_ms_per_pass": {"block_matching": s2.ms_block_matching / n2, "groups": g_ms[1], "aggregate": a_ms[1]}}

    # ---- CPU baseline on the host cores: one window pass per step of the same workload (bounded sample) ----
    cpu = None
    if rank == 0 and not args.no_cpu:
        cpu = cpu_sample(kind_hint="auto")

    if rank == 0:
        line = {"metric": "LF Mpix/s (both steps) 17x17x1024^2 RGB", "value": value, "unit": "LF Mpix/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD, "per_rank": "one full light field per GPU (replicas; no data-path collective)",
                           "l2": "inputs (3.6 GB per buffer) larger than L2", "passes_per_step": args.passes or N_PASSES},
                "e2e": e2e, "gpu_launches": launches, "clocks": sampler.summary(), "roofline": roof, "roofline_other": roof_other, "roofline_bm": roof_bm,
                "phases": phases, "strong_scaling": strong, "cpu_baseline": cpu, "psnr": {"noisy": psnr_in, "denoised": psnr_out}}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def ncu_traffic():
    """DRAM bytes per launch of the main kernels from the newest committed ncu capture (profiles/*_traffic.json, written by
    tools/ncu_kernel_table.py from an `ncu --set full` run of this same workload)."""
    import glob
    files = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "*_traffic.json")))
    if not files:
        return {}
    d = json.load(open(files[-1]))
    d["_file"] = "profiles/" + os.path.basename(files[-1])
    return d


def cpu_sample(kind_hint="auto", size=None, threads=None):
    """Time one window pass per step (3x3 SAIs = one angular window) on the host cores and extrapolate to the 17x17 LF:
    every window pass of the `pst == cst` branch costs the same (SURVEY.md 8(d) CPU baseline)."""
    import oracleapi as O
    import lfdata
    size = size or int(os.environ.get("LFBM5D_CPU_SAMPLE", "256"))
    cores = threads or os.cpu_count() or 1
    clean = lfdata.synth_lf(3, 3, size, size)
    noisy = O.add_noise(clean, CFG["sigma"])
    mask = np.ones(9, np.uint32)
    use_ref = False
    if kind_hint in ("auto", "reference"):
        try:
            import refapi as R
            use_ref = R.available()
        except Exception:
            use_ref = False
    s1, s2 = CFG["s1"], CFG["s2"]
    t0 = time.perf_counter()
    if use_ref:
        nb = 1
        while nb * 2 <= cores:
            nb *= 2                              # main.cpp:96-105: cores rounded down to a power of two
        os.environ["LFBM5D_REF_OMP_THREADS"] = str(cores)
        b, n = R.run_step1(noisy, mask, CFG["sigma"], CFG["lam"], 3, 3, 1, s1["N"], s1["nSim"], s1["nDisp"], s1["k"], s1["p"], R.ID, R.SADCT, R.HAAR, nb_threads=nb)
        R.run_step2(n, b, mask, CFG["sigma"], 3, 3, 1, s2["N"], s2["nSim"], s2["nDisp"], s2["k"], s2["p"], R.DCT, R.SADCT, R.HAAR, nb_threads=nb)
        kind, used = "reference", nb
    else:
        O.lib().orc_set_threads(cores)
        b, n, _ = O.run_step1(noisy, mask, CFG["sigma"], CFG["lam"], 3, 3, 1, s1["N"], s1["nSim"], s1["nDisp"], s1["k"], s1["p"], O.ID, O.SADCT, O.HAAR)
        O.run_step2(n, b, mask, CFG["sigma"], 3, 3, 1, s2["N"], s2["nSim"], s2["nDisp"], s2["k"], s2["p"], O.DCT, O.SADCT, O.HAAR)
        kind, used = "port", cores
    dt = time.perf_counter() - t0
    # one window pass over 9 SAIs of size^2; the full LF needs 64 passes over 9 SAIs of 1024^2 per step
    scale = N_PASSES * (CFG["H"] * CFG["W"]) / float(size * size)
    lf_pix = CFG["aw"] * CFG["ah"] * CFG["H"] * CFG["W"]
    return {"value": lf_pix / (dt * scale) / 1e6, "unit": "LF Mpix/s", "cores": used, "kind": kind, "seconds": dt,
            "sample": "one window pass per step (3x3 SAIs of %dx%d, both steps), extrapolated x%d to 64 passes of 1024^2 SAIs; DCTs via the FFTW stand-in" % (size, size, int(scale))}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    for _ in range(args.warmup):
        cpu_sample("reference", size=128)
    t0 = time.perf_counter()
    vals = [cpu_sample("reference") for _ in range(args.steps)]
    dt = time.perf_counter() - t0
    v = float(np.mean([x["value"] for x in vals]))
    cpu = dict(vals[-1])
    cpu["value"] = v
    print(json.dumps({"impl": "reference", "metric": "LF Mpix/s (both steps) 17x17x1024^2 RGB", "value": v, "unit": "LF Mpix/s",
                      "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
                      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": {"workload": WORKLOAD}, "cpu_baseline": cpu,
                      "e2e": {"value": v, "unit": "LF Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--passes", type=int, default=0, help="debug: stop each step after this many window passes (value is scaled)")
    ap.add_argument("--profile-passes", type=int, default=4)
    ap.add_argument("--no-strong", action="store_true", help="skip the one-light-field strong-scaling measurement at N > 1")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
