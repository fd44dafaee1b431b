"""Diagnostic run on a GPU box: one teacher-forced window pass per step against the oracle, with per-stage stats.
Not a test (pytest ignores it); writes gpurun_out/probe.json."""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import lfbm5d_b200 as L  # noqa: E402
import oracleapi as O  # noqa: E402
import lfdata  # noqa: E402


def stats(a, b):
    d = np.abs(a.astype(np.float64) - b.astype(np.float64))
    return dict(max=float(d.max()), p9999=float(np.quantile(d, 0.9999)), frac_gt_1e3=float((d > 1e-3).mean()),
                ref_absmax=float(np.abs(b).max()))


def main():
    H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (40, 48)
    out = {}
    clean = lfdata.synth_lf(3, 3, H, W)
    noisy = O.add_noise(clean, 25.0)
    n = 24
    y = noisy.copy()
    for st in range(9):
        O.lib().orc_color_space_transform(O.fp(y[st]), O.OPP, W, H, 3, 1)
    sym = np.stack([O.symetrize(y[st], n) for st in range(9)])
    zero = np.zeros_like(sym)
    mask, proc = np.ones(9, np.uint32), np.zeros(9, np.uint32)
    eng = L.LFBM5D(0)
    for name, step, k, N, t2 in (("step1_id", 1, 16, 8, O.ID), ("step1_bior", 1, 16, 8, O.BIOR), ("step1_dct8", 1, 8, 16, O.DCT)):
        t = time.time()
        on, od, odbg = O.run_pass(1, sym, None, zero, zero, mask, proc, 4, 3, 25.0, 2.7, 18, 6, k, N, 4, t2, O.SADCT, O.HAAR, debug=True)
        t_or = time.time() - t
        prm = L.make_params(25.0, 2.7, 3, 3, 1, W, H, 3, N, 18, 6, k, 4, t2, L.SADCT, L.HAAR)
        t = time.time()
        gn, gd, gdbg = eng.debug_pass(1, prm, sym, None, zero, zero, mask, proc, 4, debug=True)
        t_gpu = time.time() - t
        out[name] = dict(count_eq=bool(np.array_equal(odbg[0], gdbg[0])), idx_eq=bool(np.array_equal(odbg[1], gdbg[1])),
                         idx_mismatch_rows=int((odbg[1] != gdbg[1]).any(1).sum()),
                         first_eq=bool(np.array_equal(odbg[2], gdbg[2])), first_mismatch=int((odbg[2] != gdbg[2]).sum()),
                         shape_eq=bool(np.array_equal(odbg[3], gdbg[3])), num=stats(gn, on), den=stats(gd, od),
                         num_equal=bool(np.array_equal(gn, on)), t_oracle=t_or, t_gpu=t_gpu)
        print(name, json.dumps(out[name]))
        if name == "step1_id":
            basic_sym = np.where(od != 0, on / np.where(od != 0, od, 1), sym).astype(np.float32)
    for name, t5 in (("step2_dct_haar", O.HAAR), ("step2_dct_hw", O.HADAMARD)):
        on, od, odbg = O.run_pass(2, sym, basic_sym, zero, zero, mask, proc, 4, 3, 25.0, 0.0, 18, 6, 8, 16, 4, O.DCT, O.SADCT, t5, debug=True)
        prm = L.make_params(25.0, 0.0, 3, 3, 1, W, H, 3, 16, 18, 6, 8, 4, L.DCT, L.SADCT, t5)
        gn, gd, gdbg = eng.debug_pass(2, prm, sym, basic_sym, zero, zero, mask, proc, 4, debug=True)
        out[name] = dict(count_eq=bool(np.array_equal(odbg[0], gdbg[0])), idx_eq=bool(np.array_equal(odbg[1], gdbg[1])),
                         first_eq=bool(np.array_equal(odbg[2], gdbg[2])), shape_eq=bool(np.array_equal(odbg[3], gdbg[3])),
                         num=stats(gn, on), den=stats(gd, od))
        print(name, json.dumps(out[name]))
    # full steps through the host API on a 3x3 LF
    t = time.time()
    ob, onz, osch = O.run_step1(noisy, np.ones(9), 25.0, 2.7, 3, 3, 1, 8, 18, 6, 16, 4, O.ID, O.SADCT, O.HAAR)
    od2, _, _, _ = O.run_step2(onz, ob, np.ones(9), 25.0, 3, 3, 1, 16, 18, 6, 8, 4, O.DCT, O.SADCT, O.HAAR)
    t_or = time.time() - t
    p1 = L.make_params(25.0, 2.7, 3, 3, 1, W, H, 3, 8, 18, 6, 16, 4, L.ID, L.SADCT, L.HAAR)
    p2 = L.make_params(25.0, 0.0, 3, 3, 1, W, H, 3, 16, 18, 6, 8, 4, L.DCT, L.SADCT, L.HAAR)
    t = time.time()
    gb, gnz = eng.step1(p1, noisy, np.ones(9))
    gd2, _, _ = eng.step2(p2, gnz, gb, np.ones(9))
    t_gpu = time.time() - t
    out["full"] = dict(basic=stats(gb, ob), noisy_rt=stats(gnz, onz), denoised=stats(gd2, od2),
                       psnr_oracle=[O.psnr(ob, clean)[0], O.psnr(od2, clean)[0]], psnr_gpu=[O.psnr(gb, clean)[0], O.psnr(gd2, clean)[0]],
                       t_oracle=t_or, t_gpu=t_gpu, sched=eng.schedule().tolist(), launches=int(eng.stats().kernel_launches))
    print("full", json.dumps(out["full"]))
    os.makedirs(os.path.join(os.path.dirname(HERE), "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(os.path.dirname(HERE), "gpurun_out", "probe.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
