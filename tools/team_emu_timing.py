#!/usr/bin/env python
"""Per-phase device time of the team path for rank 0's share, emulated on ONE GPU (world ranks as contexts on one device, run one
after the other): how the block-matching launches behave when a rank only holds 1/world of the offset planes.
  python tools/team_emu_timing.py 1 2 4"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main():
    import torch
    import lfbm5d_b200 as L
    import run_config as RC
    dev = torch.device("cuda", 0)
    aw = ah = 3
    H = W = 1024
    clean, noisy = RC.synth(torch, dev, aw, ah, H, W, 10.0)
    mask = np.ones(9, np.uint32)
    p1 = L.make_params(10.0, 2.7, aw, ah, 1, W, H, 3, 8, 18, 6, 16, 4, L.ID, L.SADCT, L.HAAR)
    p2 = L.make_params(10.0, 0.0, aw, ah, 1, W, H, 3, 16, 18, 6, 8, 4, L.DCT, L.SADCT, L.HAAR)
    for world in [int(a) for a in sys.argv[1:]] or [1, 2, 4]:
        team = L.Team.emulated(0, world)
        ws = [noisy.clone() for _ in range(world)]
        bs = [torch.zeros_like(noisy) for _ in range(world)]
        outs = [torch.zeros_like(noisy) for _ in range(world)]
        res = {}
        for it in range(2):
            for w in ws:
                w.copy_(noisy)
            if it == 1:
                team.timing(True)
            team.step(1, p1, [t.data_ptr() for t in ws], None, mask, [t.data_ptr() for t in bs], gather=0)
            if it == 1:
                res["step1"] = team.timing(False)
                team.timing(True)
            team.step(2, p2, [t.data_ptr() for t in ws], [t.data_ptr() for t in bs], mask, [t.data_ptr() for t in outs], gather=0)
            if it == 1:
                res["step2"] = team.timing(False)
        names = ["pad", "x_est0", "bm", "x_match", "sel_groups_agg1", "x_border", "agg2", "x_border_back", "bm_self_planes", "bm_partial", "bm_stereo_planes", "bm_argmin"]
        print(json.dumps({"world": world, "ms_per_pass_rank0": {k: dict(zip(names, [round(x, 3) for x in v])) for k, v in res.items()}}), flush=True)
        team.close()
        del ws, bs, outs
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
