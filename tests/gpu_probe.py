"""Scratch probe run on the GPU box while debugging (not a test)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import oracleapi as O, lfdata
import lfbm5d_b200 as L
eng = L.LFBM5D(0)
for (aw, H, W, masked) in [(3, 28, 32, False), (5, 24, 28, False)]:
    clean = lfdata.synth_lf(aw, aw, H, W)[:, :1]
    noisy = O.add_noise(np.ascontiguousarray(clean), 25.0)
    m = np.ones(aw * aw, np.uint32)
    ob, onrt, sched = O.run_step1(noisy, m, 25.0, 2.7, aw, aw, 1, 8, 18, 6, 16, 4, O.ID, O.SADCT, O.HAAR)
    p2_ = L.make_params(25.0, 0.0, aw, aw, 1, W, H, 1, 16, 18, 6, 8, 4, L.DCT, L.SADCT, L.HAAR)
    for mp in (1, 2, 0):
        eng.set_max_passes(mp)
        d, _, _ = eng.step2(p2_, noisy, ob, m)
        gs = eng.schedule().copy()
        od, _, _, s2 = O.run_step2(noisy, ob, m, 25.0, aw, aw, 1, 16, 18, 6, 8, 4, O.DCT, O.SADCT, O.HAAR, max_passes=mp)
        diff = np.abs(d - od)
        print(aw, "max_passes", mp, "sched gpu", gs.tolist(), "oracle", s2.tolist())
        print("   frac>1e-3 per SAI", [(round(float((diff[i] > 1e-3).mean()), 3)) for i in range(aw * aw)], "max", diff.max())
