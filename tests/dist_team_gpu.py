"""One light field on N GPUs with the team path (csrc/team.cuh) over NCCL (torchrun, one rank per GPU): bit-identical to the
single-GPU run, with timing.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29577 tests/dist_team_gpu.py [--big] [--no-check]"""
import json
import os
import sys
import time

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")      # lanes: concurrent streams must not share a hardware queue

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main():
    import torch
    import torch.distributed as dist
    import lfbm5d_b200 as L
    from lfbm5d_b200 import dist as D
    import run_config as RC
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    big = "--big" in sys.argv
    check = "--no-check" not in sys.argv
    aw, ah, H, W = (17, 17, 1024, 1024) if big else (4, 3, 96 * world + 160, 80)
    eng = L.LFBM5D(local)
    clean, noisy = RC.synth(torch, dev, aw, ah, H, W, 10.0)          # same seed on every rank: identical replicas
    mask = np.ones(aw * ah, np.uint32)
    if not big:
        mask[5] = 0
    p1 = L.make_params(10.0, 2.7, aw, ah, 1, W, H, 3, 8, 18, 6, 16, 4, L.ID, L.DCT if not big else L.SADCT, L.HAAR)
    p2 = L.make_params(10.0, 0.0, aw, ah, 1, W, H, 3, 16, 18, 6, 8, 4, L.DCT, L.DCT if not big else L.SADCT, L.HAAR)
    team = D.make_team(eng, dist, dev)
    lanes = [int(a.split("=")[1]) for a in sys.argv if a.startswith("--lanes=")]
    if lanes and lanes[0] > 1:
        team.set_lanes(lanes[0])
    stream = torch.cuda.ExternalStream(eng.stream(), device=dev)
    times = {}
    work, basic, out = noisy.clone(), torch.zeros_like(noisy), torch.zeros_like(noisy)
    for it in range(3 if big else 1):                                  # big: one warm-up (allocations, NCCL connections), one timed, one with phase events
        work.copy_(noisy)
        if it == 2:
            t_keep = times["team"]
            team.timing(True)
        torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
        team.step(1, p1, [work.data_ptr()], None, mask, [basic.data_ptr()], gather=0)
        team.step(2, p2, [work.data_ptr()], [basic.data_ptr()], mask, [out.data_ptr()], gather=1)
        torch.cuda.synchronize(); dist.barrier(); times["team"] = time.perf_counter() - t0
    phases = None
    if big:
        phases = team.timing(False)
        times["team"] = t_keep
    res_team = (out.clone(), work.clone())
    st = team.stats()
    band = team.band(rank)
    ok = [1, 1]
    if check:
        work, basic, out = noisy.clone(), torch.zeros_like(noisy), torch.zeros_like(noisy)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        eng.step1_device(p1, work.data_ptr(), mask, basic.data_ptr())
        eng.step2_device(p2, work.data_ptr(), basic.data_ptr(), mask, out.data_ptr())
        torch.cuda.synchronize(); times["single"] = time.perf_counter() - t0
        ok = [int(torch.equal(res_team[0], out)), int(torch.equal(res_team[1], work))]
    flags = torch.tensor(ok, device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    tt = torch.tensor([times["team"], times.get("single", 0.0)], device=dev)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"world": world, "lf": [ah, aw, H, W], "team_s": float(tt[0]), "single_s": float(tt[1]),
                          "speedup": float(tt[1] / tt[0]) if check else None, "denoised_identical": bool(flags[0]), "noisy_roundtrip_identical": bool(flags[1]),
                          "band_rank0": band, "bytes_sent_rank0": st["bytes_exchanged"], "passes_redone": st["passes_redone"],
                          "tie_patches_rank0": st["tie_patches"], "peer_view": st["peer_view"],
                          "phase_ms_rank0": phases, "lanes": lanes[0] if lanes else 1}), flush=True)
    team.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flags.min()) == 1 else 1)


if __name__ == "__main__":
    main()
