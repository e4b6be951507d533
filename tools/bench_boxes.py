#!/usr/bin/env python3
"""Small-box path: PeleC hands the reactor 16^3 - 64^3 boxes (amr.max_grid_size / blocking_factor).  Times
nbox boxes of n^3 cells (a) one prx_react_box call after the other on one stream - the reference's MFIter loop with its
streamSynchronize per box (PC/Source/React.cpp:133-230) - and (b) one prx_react_boxes call.  Prints one JSON line per case.
usage: python tools/bench_boxes.py [--mech LiDryer] [--reactor ReactorCvode]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mech", default="LiDryer")
    ap.add_argument("--reactor", default="ReactorCvode")
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    import numpy as np
    import torch
    import bench
    from pelec_b200.reactor import Array4, Box, ReactorB200
    dev = torch.device("cuda:0")
    parm = {"ode__atol": 1e-10}
    if a.reactor == "ReactorCvode":
        parm.update({"ode__rtol": 1e-10, "cvode__max_order": 4, "cvode__solve_type": "denseAJ_direct"})
    r = ReactorB200.create(a.reactor, a.mech, **parm)
    r.init(1, 1)
    stream = torch.cuda.current_stream(dev)
    for n, nbox in ((16, 64), (32, 64), (64, 8)):
        rYj, Tj, rEj = bench.profile_states(a.mech, 128)  # the 128 distinct ReactEval states; box b takes a j-window of them
        N = rYj.shape[0]
        pristine = []
        for b in range(nbox):
            j0 = (b * 7) % (128 - n + 1)
            ex = lambda v, nc: torch.from_numpy(np.ascontiguousarray(v)).to(dev).reshape(nc, 1, n, 1).expand(nc, n, n, n).contiguous()
            pristine.append((ex(rYj[:, j0:j0 + n], N), ex(Tj[j0:j0 + n], 1), ex(rEj[j0:j0 + n], 1)))
        zsrc = [torch.zeros((N, n, n, n), dtype=torch.float64, device=dev) for _ in range(nbox)]
        zes = [torch.zeros((1, n, n, n), dtype=torch.float64, device=dev) for _ in range(nbox)]
        box = Box((0, 0, 0), (n - 1, n - 1, n - 1))

        def fresh():
            return [[p[0].clone(), zsrc[i], p[1].clone(), p[2].clone(), zes[i], torch.zeros((1, n, n, n), dtype=torch.float64, device=dev)]
                    for i, p in enumerate(pristine)]
        res = {}
        for mode in ("box_by_box_sync", "box_by_box_async", "react_boxes"):
            times = []
            for rep in range(a.reps + 1):
                arrs = fresh()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                if mode == "react_boxes":
                    r.react_boxes([box] * nbox, *[[Array4(x[i]) for x in arrs] for i in range(6)], None, 1e-6, 0.0, stream=stream.cuda_stream)
                else:
                    for x in arrs:
                        r.react(box, *[Array4(t) for t in x], None, 1e-6, 0.0, stream=stream.cuda_stream)
                        if mode == "box_by_box_sync":
                            torch.cuda.synchronize()  # React.cpp:230
                e1.record(stream)
                torch.cuda.synchronize()
                if rep > 0:
                    times.append(e0.elapsed_time(e1))
            ms = sorted(times)[len(times) // 2]
            res[mode] = {"ms": ms, "cells_per_s": nbox * n ** 3 / (ms * 1e-3)}
        print(json.dumps({"mechanism": a.mech, "reactor": a.reactor, "box": "%d^3" % n, "boxes": nbox, "cells": nbox * n ** 3, **res}))
    r.close()


if __name__ == "__main__":
    main()
