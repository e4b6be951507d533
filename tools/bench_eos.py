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
    ap.add_argument("--kernel", default="compute_temp", choices=["compute_temp", "ctoprim", "transport"],
                    help="ctoprim: pc_ctoprim, (7+N) doubles in, (8+N) + 6 out; transport: get_transport_coeffs, N+2 in, N+3 out")
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
    if a.kernel != "compute_temp":
        return bench_prim(a, r, S, box, stream, N, n, dev)
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


def bench_prim(a, r, S, box, stream, N, n, dev):
    import torch
    from pelec_b200.reactor import Array4
    S[1:4] = 3.0e3 * S[0]
    S[4] = S[5] + 0.5 * 3 * 9.0e6 * S[0]
    q = torch.zeros((8 + N, n, n, n), dtype=torch.float64, device=dev)
    qa = torch.zeros((6, n, n, n), dtype=torch.float64, device=dev)
    coef = torch.zeros((N + 3, n, n, n), dtype=torch.float64, device=dev)
    r.ctoprim([box], [Array4(S)], [Array4(q)], [Array4(qa)], stream=stream.cuda_stream)

    def call():
        if a.kernel == "ctoprim":
            r.ctoprim([box], [Array4(S)], [Array4(q)], [Array4(qa)], stream=stream.cuda_stream)
        else:
            r.transport([box], [Array4(q[8:])], [Array4(q[7:8])], [Array4(q[0:1])], [Array4(coef[:N])], [Array4(coef[N:N + 1])],
                        [Array4(coef[N + 1:N + 2])], [Array4(coef[N + 2:N + 3])], stream=stream.cuda_stream)
    ms = []
    for rep in range(8):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        call()
        e1.record(stream)
        torch.cuda.synchronize()
        if rep >= 3:
            ms.append(e0.elapsed_time(e1))
    t = sorted(ms)[len(ms) // 2]
    nbytes = ((21 + 2 * N) if a.kernel == "ctoprim" else (2 * N + 5)) * 8 * n ** 3
    peak = 6544.3
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    print(json.dumps({"kernel": a.kernel + "_kernel_" + a.mech, "cells": n ** 3, "ms": t, "algorithmic_bytes": nbytes,
                      "achieved_GBps": nbytes / (t * 1e-3) / 1e9, "peak_GBps": peak, "frac": nbytes / (t * 1e-3) / 1e9 / peak,
                      "cells_per_s": n ** 3 / (t * 1e-3), "note": "FABs > L2; timed call after 3 warm-up calls"}))
    r.close()


if __name__ == "__main__":
    main()
