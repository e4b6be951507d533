#!/usr/bin/env python3
"""HBM-bound side of the path: prx_compute_temp_boxes (pc_cmpTemp, PC/Source/Utilities.H:70-84) on a 256^3 state FAB.
Algorithmic bytes per cell = (N + 3) doubles read (rho, rho*e, T guess, rhoY_k) + 1 written.  Prints one JSON line with the
achieved GB/s against MEASURED_PEAKS.json hbm_gbs.  usage: python tools/bench_eos.py [--mech LiDryer] [--n 256]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mech", default="LiDryer")
    ap.add_argument("--n", type=int, default=256)
    a = ap.parse_args()
    import numpy as np
    import torch
    import bench
    from pelec_b200.reactor import Array4, Box, ReactorB200
    dev = torch.device("cuda:0")
    n = a.n
    rYj, Tj, rEj = bench.profile_states(a.mech, n)
    N = rYj.shape[0]
    S = torch.zeros((7 + N, n, n, n), dtype=torch.float64, device=dev)
    ex = lambda v, nc: torch.from_numpy(np.ascontiguousarray(v)).to(dev).reshape(nc, 1, n, 1).expand(nc, n, n, n)
    S[7:] = ex(rYj, N)
    S[0] = S[7:].sum(dim=0)
    S[5] = ex(rEj, 1)[0]
    S[6] = ex(Tj, 1)[0] + 25.0
    r = ReactorB200.create("ReactorRK64", a.mech)
    r.init(1, 1)
    box = Box((0, 0, 0), (n - 1, n - 1, n - 1))
    stream = torch.cuda.current_stream(dev)
    ms = []
    for rep in range(8):
        S[6] = ex(Tj, 1)[0] + 25.0
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        r.compute_temp([box], [Array4(S)], stream=stream.cuda_stream)
        e1.record(stream)
        torch.cuda.synchronize()
        if rep >= 3:
            ms.append(e0.elapsed_time(e1))
    err = float((S[6] - ex(Tj, 1)[0]).abs().max())
    t = sorted(ms)[len(ms) // 2]
    nbytes = (N + 4) * 8 * n ** 3
    peak = 6544.3
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    print(json.dumps({"kernel": "compute_temp_kernel_" + a.mech, "cells": n ** 3, "ms": t, "algorithmic_bytes": nbytes,
                      "achieved_GBps": nbytes / (t * 1e-3) / 1e9, "peak_GBps": peak, "frac": nbytes / (t * 1e-3) / 1e9 / peak,
                      "cells_per_s": n ** 3 / (t * 1e-3), "max_abs_dT_vs_input_profile": err,
                      "note": "state FAB of %.1f GB > L2" % ((7 + N) * 8 * n ** 3 / 1e9)}))
    r.close()


if __name__ == "__main__":
    main()
