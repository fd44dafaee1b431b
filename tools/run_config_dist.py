#!/usr/bin/env python
"""The other BASELINE.json configurations on N GPUs (torchrun, one rank per GPU); one JSON line per run on rank 0.
  torchrun ... tools/run_config_dist.py 4            # LFBM3D per SAI on 17x17x1024^2 (`16 16 8 3 bior / 32 16 8 3 dct`): SAIs sharded over the ranks,
                                                     #   no collective in the compute path (dist.shard_mask)
  torchrun ... tools/run_config_dist.py 5 [--sigmas 10,25,50]   # 9x9 SAIs 2048x2048, README parameters: ONE light field on the team path
  torchrun ... tools/run_config_dist.py 2            # EPFL-shaped 15x15 SAIs 434x625, `1 18 3 16 3 bior sadct haar / 8 18 3 8 3 dct sadct haar`: team path
Inputs: synthetic clean light field (tests/lfdata.py) + the reference's host noise (mt19937ar, seed 20171016 + st), built on rank 0 and
broadcast. Timing: CUDA events on the library's stream, max over ranks. These are parity-test shapes, not bench lines."""
import argparse
import json
import os
import sys

import numpy as np

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")      # team lanes: concurrent streams must not share a hardware queue

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import lfbm5d_b200 as L
    from lfbm5d_b200 import dist as D
    import lfdata
    ap = argparse.ArgumentParser()
    ap.add_argument("config", type=int, choices=(2, 4, 5))
    ap.add_argument("--sigmas", type=str, default="")
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--lanes", type=int, default=1)
    ap.add_argument("--lf", type=str, default="", help="aw,ah,H,W instead of the configuration's light field (small shapes: more ranks than row bands)")
    ap.add_argument("--no-peer", action="store_true", help="team exchanges through NCCL send / recv instead of the peer-memory kernels")
    ap.add_argument("--check", action="store_true", help="compare the team's result with the single-GPU run on every rank (bit for bit)")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    eng = L.LFBM5D(local)
    stream = torch.cuda.ExternalStream(eng.stream(), device=dev)
    team = D.make_team(eng, dist, dev) if world > 1 and args.config != 4 else None
    if team is not None and args.lanes > 1:
        team.set_lanes(args.lanes)
    if team is not None and args.no_peer:
        team.use_peer_exchange(False)
    shapes = {2: (15, 15, 434, 625), 4: (17, 17, 1024, 1024), 5: (9, 9, 2048, 2048)}
    aw, ah, H, W = shapes[args.config] if not args.lf else tuple(int(x) for x in args.lf.split(","))
    asize = aw * ah
    sigmas = [float(s) for s in args.sigmas.split(",")] if args.sigmas else ([10.0, 25.0, 50.0] if args.config == 5 else [10.0])
    clean = torch.empty((asize, 3, H, W), device=dev)
    if rank == 0:
        clean_h = lfdata.synth_lf(aw, ah, H, W)
        clean.copy_(torch.from_numpy(clean_h))
    if world > 1:
        dist.broadcast(clean, src=0)
    mask = np.ones(asize, np.uint32)

    def timed(fn):
        best = None
        for _ in range(args.reps):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            e1.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = float(t.item()) if best is None else min(best, float(t.item()))
        return best

    for sigma in sigmas:
        noisy = torch.empty_like(clean)
        if rank == 0:
            noisy.copy_(torch.from_numpy(L.add_noise(clean.cpu().numpy(), sigma)))
        if world > 1:
            dist.broadcast(noisy, src=0)
        work, basic, out = torch.empty_like(noisy), torch.zeros_like(noisy), torch.zeros_like(noisy)
        info = {}
        if args.config == 4:
            prm = L.make_params3d(sigma, asize, W, H, 3, 16, 16, 8, 8, 16, 32, 3, 3, L.BIOR, L.DCT, 2.7)
            m = D.shard_mask(mask, world, rank) if world > 1 else mask
            lo, hi = D.shard_sais(asize, world, rank) if world > 1 else (0, asize)

            def run():
                with torch.cuda.stream(stream):
                    work.copy_(noisy, non_blocking=True)
                eng.bm3d_device(prm, work.data_ptr(), m, basic.data_ptr(), out.data_ptr())
            ms = timed(run)
            sq = torch.tensor([float(((out[lo:hi] - clean[lo:hi]) ** 2).sum().item()), float((hi - lo) * 3 * H * W)], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(sq)
            psnr = float(10.0 * np.log10(255.0 ** 2 / (sq[0].item() / sq[1].item())))
            info = {"config": 4, "path": "LFBM3D (run_bm3d_LF), 16 16 8 3 bior / 32 16 8 3 dct, SAIs sharded over the ranks, no data-path collective",
                    "sais_per_rank": [hi - lo for lo, hi in [D.shard_sais(asize, world, r) for r in range(world)]] if world > 1 else [asize]}
        else:
            if args.config == 2:
                p1 = L.make_params(sigma, 2.7, aw, ah, 1, W, H, 3, 1, 18, 3, 16, 3, L.BIOR, L.SADCT, L.HAAR)
                p2 = L.make_params(sigma, 0.0, aw, ah, 1, W, H, 3, 8, 18, 3, 8, 3, L.DCT, L.SADCT, L.HAAR)
            else:
                p1 = L.make_params(sigma, 2.7, aw, ah, 1, W, H, 3, 8, 18, 6, 16, 4, L.ID, L.SADCT, L.HAAR)
                p2 = L.make_params(sigma, 0.0, aw, ah, 1, W, H, 3, 16, 18, 6, 8, 4, L.DCT, L.SADCT, L.HAAR)

            def run():
                with torch.cuda.stream(stream):
                    work.copy_(noisy, non_blocking=True)
                if team is None:
                    eng.step1_device(p1, work.data_ptr(), mask, basic.data_ptr())
                    eng.step2_device(p2, work.data_ptr(), basic.data_ptr(), mask, out.data_ptr())
                else:
                    team.step(1, p1, [work.data_ptr()], None, mask, [basic.data_ptr()], gather=0)
                    team.step(2, p2, [work.data_ptr()], [basic.data_ptr()], mask, [out.data_ptr()], gather=1)
            ms = timed(run)
            psnr = float(10.0 * torch.log10(255.0 ** 2 / torch.mean((out - clean) ** 2)))
            sched = eng.schedule()
            info = {"config": args.config, "path": "LFBM5D, one light field on %d GPU(s)%s" % (world, " (team path: every window pass split inside the window)" if team else ""),
                    "windows_step2": int(len(sched)), "core_calls_step2": int(sched[:, 3].sum()) if len(sched) else 0}
            if team is not None:
                info["team"] = team.stats()
                info["team"]["peer_exchange"] = not args.no_peer
            if team is not None and args.check:
                res = out.clone()
                with torch.cuda.stream(stream):
                    work.copy_(noisy, non_blocking=True)
                eng.step1_device(p1, work.data_ptr(), mask, basic.data_ptr())
                eng.step2_device(p2, work.data_ptr(), basic.data_ptr(), mask, out.data_ptr())
                torch.cuda.synchronize()
                same = torch.tensor([int(torch.equal(res, out))], device=dev)
                dist.all_reduce(same, op=dist.ReduceOp.MIN)
                info["identical_to_single_gpu"] = bool(int(same.item()))
                info["max_abs_diff_rank0"] = float((res - out).abs().max())
                info["psnr_single_gpu"] = float(10.0 * torch.log10(255.0 ** 2 / torch.mean((out - clean) ** 2)))
        if rank == 0:
            info.update({"n_gpus": world, "lf": [ah, aw, H, W], "sigma": sigma, "ms": ms, "lf_mpix_per_s": asize * H * W / (ms * 1e-3) / 1e6,
                         "psnr_noisy": float(10.0 * torch.log10(255.0 ** 2 / torch.mean((noisy - clean) ** 2))), "psnr_denoised": psnr})
            print(json.dumps(info), flush=True)
        del noisy, work, basic, out
        torch.cuda.empty_cache()
    if team is not None:
        team.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
