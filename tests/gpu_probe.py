"""Scratch probe run on the GPU box while debugging (not a test)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tools"))
import torch
import lfbm5d_b200 as L
import run_config as RC
dev = torch.device("cuda", 0)
eng = L.LFBM5D(0)
aw = ah = 9; H = W = 2048
clean, noisy = RC.synth(torch, dev, aw, ah, H, W, 10.0)
work, basic, out = noisy.clone(), torch.empty_like(noisy), torch.empty_like(noisy)
mask = np.ones(aw * ah, np.uint32)
p1 = L.make_params(10.0, 2.7, aw, ah, 1, W, H, 3, 8, 18, 6, 16, 4, L.ID, L.SADCT, L.HAAR)
p2 = L.make_params(10.0, 0.0, aw, ah, 1, W, H, 3, 16, 18, 6, 8, 4, L.DCT, L.SADCT, L.HAAR)
for rep in range(2):
    eng.set_max_passes(2)
    eng.enable_timing(True)
    for step in (1, 2):
        eng.reset_stats()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        if step == 1: eng.step1_device(p1, work.data_ptr(), mask, basic.data_ptr())
        else: eng.step2_device(p2, work.data_ptr(), basic.data_ptr(), mask, out.data_ptr())
        torch.cuda.synchronize(); t1 = time.perf_counter()
        s = eng.stats()
        print("rep", rep, "step", step, "wall %.3f" % (t1 - t0), "passes", s.window_passes, "bm %.1f sat %.1f groups %.1f agg %.1f other %.1f" % (s.ms_block_matching / s.window_passes, s.ms_sat / s.window_passes, s.ms_groups / s.window_passes, s.ms_aggregate / s.window_passes, s.ms_other / s.window_passes), flush=True)
