"""Window-level multi-GPU run of one light field (torchrun, one rank per GPU): bit-identical to the single-GPU run.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29577 tests/dist_windows_gpu.py [--big]"""
import os
import sys
import time

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
    aw, ah, H, W = (17, 17, 1024, 1024) if big else (7, 7, 72, 80)
    eng = L.LFBM5D(local)
    clean, noisy = RC.synth(torch, dev, aw, ah, H, W, 10.0)
    mask = np.ones(aw * ah, np.uint32)
    if not big:
        mask[3] = 0                                    # an empty SAI: sticky sadct switch (bm5d.cpp:276-280)
    p1 = L.make_params(10.0, 2.7, aw, ah, 1, W, H, 3, 8, 18, 6, 16, 4, L.ID, L.DCT if not big else L.SADCT, L.HAAR)
    p2 = L.make_params(10.0, 0.0, aw, ah, 1, W, H, 3, 16, 18, 6, 8, 4, L.DCT, L.DCT if not big else L.SADCT, L.HAAR)
    res = {}
    for mode in ("windows", "sequential"):
        work, basic, out = noisy.clone(), torch.zeros_like(noisy), torch.zeros_like(noisy)
        torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
        if mode == "windows":
            plan = D.run_step_windows(eng, 1, p1, work.data_ptr(), 0, mask, basic.data_ptr(), dist, dev)
            D.run_step_windows(eng, 2, p2, work.data_ptr(), basic.data_ptr(), mask, out.data_ptr(), dist, dev)
        else:
            eng.step1_device(p1, work.data_ptr(), mask, basic.data_ptr())
            seq_sched = eng.schedule().copy()
            eng.step2_device(p2, work.data_ptr(), basic.data_ptr(), mask, out.data_ptr())
        torch.cuda.synchronize(); dist.barrier(); t1 = time.perf_counter()
        res[mode] = (basic.clone(), out.clone(), t1 - t0)
    same_b = bool(torch.equal(res["windows"][0], res["sequential"][0]))
    same_d = bool(torch.equal(res["windows"][1], res["sequential"][1]))
    # the static plan is the schedule the sequential driver takes
    plan_ok = len(plan) == len(seq_sched) and all(int(a[2]) == int(b[1]) and int(a[3]) == int(b[2]) for a, b in zip(plan, seq_sched))
    flags = torch.tensor([int(same_b), int(same_d), int(plan_ok)], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("world %d %dx%dx%dx%d: windows %.3f s, sequential (one GPU) %.3f s, speed-up %.2f; levels %d; basic identical %s, denoised identical %s, plan == schedule %s"
              % (world, ah, aw, H, W, res["windows"][2], res["sequential"][2], res["sequential"][2] / res["windows"][2], int(plan[:, 4].max()) + 1,
                 bool(flags[0]), bool(flags[1]), bool(flags[2])), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flags.min()) == 1 else 1)


if __name__ == "__main__":
    main()
