#!/usr/bin/env python
"""Run the other BASELINE.json configurations once on one GPU (device-resident API) and report throughput and PSNR.
  python tools/run_config.py 2      # EPFL-shaped 15x15 SAIs 434x625, `1 18 3 16 3 bior sadct haar / 8 18 3 8 3 dct sadct haar`
  python tools/run_config.py 4      # LFBM3D per SAI on 17x17x1024^2, `16 16 8 3 bior / 32 16 8 3 dct` (--sais N: first N SAIs only)
  python tools/run_config.py 5      # 9x9 SAIs 2048x2048, README parameters, sigma 10 / 25 / 50
These are parity-test shapes, not bench lines (bench.py measures configs[2])."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def synth(torch, dev, aw, ah, H, W, sigma, seed=7):
    import lfdata
    pad = max(aw, ah)
    base = torch.from_numpy(lfdata.base_image(H + 2 * pad, W + 2 * pad, 3, seed=seed)).to(dev)
    clean = torch.empty((aw * ah, 3, H, W), device=dev)
    for s_ in range(ah):
        for t_ in range(aw):
            oy, ox = pad + (s_ - ah // 2), pad + (t_ - aw // 2)
            clean[s_ * aw + t_] = base[:, oy:oy + H, ox:ox + W]
    gen = torch.Generator(device=dev)
    gen.manual_seed(20171016)
    return clean, clean + sigma * torch.randn(clean.shape, device=dev, generator=gen)


def psnr(torch, a, b):
    return float(10.0 * torch.log10(255.0 ** 2 / torch.mean((a - b) ** 2)))


def main():
    import torch
    import lfbm5d_b200 as L
    ap = argparse.ArgumentParser()
    ap.add_argument("config", type=int, choices=(2, 4, 5))
    ap.add_argument("--sais", type=int, default=0)
    ap.add_argument("--sigmas", type=str, default="")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    eng = L.LFBM5D(0)
    res = []
    if args.config in (2, 5):
        if args.config == 2:
            aw, ah, H, W, sigmas = 15, 15, 434, 625, [10.0]
            s1 = (1, 18, 3, 16, 3, L.BIOR)
            s2 = (8, 18, 3, 8, 3, L.DCT)
        else:
            aw, ah, H, W, sigmas = 9, 9, 2048, 2048, [10.0, 25.0, 50.0]
            s1 = (8, 18, 6, 16, 4, L.ID)
            s2 = (16, 18, 6, 8, 4, L.DCT)
        if args.sigmas:
            sigmas = [float(x) for x in args.sigmas.split(",")]
        mask = np.ones(aw * ah, np.uint32)
        for sigma in sigmas:
            clean, noisy = synth(torch, dev, aw, ah, H, W, sigma)
            work, basic, out = noisy.clone(), torch.empty_like(noisy), torch.empty_like(noisy)
            p1 = L.make_params(sigma, 2.7, aw, ah, 1, W, H, 3, *s1[:5], s1[5], L.SADCT, L.HAAR)
            p2 = L.make_params(sigma, 0.0, aw, ah, 1, W, H, 3, *s2[:5], s2[5], L.SADCT, L.HAAR)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            eng.step1_device(p1, work.data_ptr(), mask, basic.data_ptr())
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            n1 = len(eng.schedule())
            eng.step2_device(p2, work.data_ptr(), basic.data_ptr(), mask, out.data_ptr())
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            res.append({"config": args.config, "sigma": sigma, "shape": [ah, aw, H, W], "passes": [n1, len(eng.schedule())],
                        "s_step1": t1 - t0, "s_step2": t2 - t1, "lf_mpix_s": aw * ah * H * W / (t2 - t0) / 1e6,
                        "psnr_noisy": psnr(torch, noisy, clean), "psnr_basic": psnr(torch, basic, clean), "psnr_denoised": psnr(torch, out, clean)})
            print(json.dumps(res[-1]), flush=True)
            del clean, noisy, work, basic, out
    else:
        aw = ah = 17
        H = W = 1024
        n = args.sais or aw * ah
        clean, noisy = synth(torch, dev, aw, ah, H, W, 10.0)
        clean, noisy = clean[:n].contiguous(), noisy[:n].contiguous()
        work, basic, out = noisy.clone(), torch.empty_like(noisy), torch.empty_like(noisy)
        p3 = L.make_params3d(10.0, n, W, H, 3, 16, 16, 8, 8, 16, 32, 3, 3, L.BIOR, L.DCT)
        mask = np.ones(n, np.uint32)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        eng.bm3d_device(p3, work.data_ptr(), mask, basic.data_ptr(), out.data_ptr())
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        res.append({"config": 4, "sais": n, "s_total": t1 - t0, "lf_mpix_s": n * H * W / (t1 - t0) / 1e6,
                    "psnr_noisy": psnr(torch, noisy, clean), "psnr_basic": psnr(torch, basic, clean), "psnr_denoised": psnr(torch, out, clean)})
        print(json.dumps(res[-1]), flush=True)


if __name__ == "__main__":
    main()
