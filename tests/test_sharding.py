"""Multi-GPU host logic on CPU: box chopping, knapsack ownership and the single statistics all-reduce,
run as 2 processes over gloo (the GPU path uses the same code over NCCL).  The oracle stands in for the
device here (test infrastructure): what is under test is that sharding reproduces the unsharded result
cell for cell and that the reduced statistics equal the unsharded totals."""
import os
import socket

import numpy as np
import pytest

from pelec_b200 import sharding

torch = pytest.importorskip("torch")


def test_chop_and_knapsack():
    boxes = sharding.chop_domain((256, 256, 256), 64)
    assert len(boxes) == 64 and sum(map(sharding.box_cells, boxes)) == 256 ** 3
    assert boxes[0] == ((0, 0, 0), (63, 63, 63)) and boxes[1][0] == (64, 0, 0)  # x fastest
    # BoxList::maxSize halves chunk and length while both are even (AMReX_BoxList.cpp:773-785): 10 / 4 -> 4, 4, 2 ; 7 / 4 -> 4, 3
    ragged = sharding.chop_domain((10, 7, 1), 4)
    assert [b[1][0] - b[0][0] + 1 for b in ragged[:3]] == [4, 4, 2]
    assert [b[1][0] - b[0][0] + 1 for b in sharding.chop_domain((100, 1, 1), 32)] == [28, 24, 24, 24]
    assert [b[1][0] - b[0][0] + 1 for b in sharding.chop_domain((96, 1, 1), 64)] == [64, 32]
    seen = np.zeros((10, 7, 1), dtype=int)
    for lo, hi in ragged:
        seen[lo[0]:hi[0] + 1, lo[1]:hi[1] + 1, lo[2]:hi[2] + 1] += 1
    assert np.all(seen == 1)
    for n in (1, 2, 4, 8):
        own = sharding.knapsack([sharding.box_cells(b) for b in boxes], n)
        assert sorted(set(own)) == list(range(n)) and all(own.count(r) == 64 // n for r in range(n))
    # cost-weighted: one expensive box gets a rank to itself
    own = sharding.knapsack([100, 1, 1, 1, 1], 2)
    assert own[0] != own[1] and len({own[i] for i in range(1, 5)}) == 1
    assert sharding.reduce_stats({"cells": 3, "max_steps": 7}) == {"cells": 3, "max_steps": 7}
    with pytest.raises(ValueError):
        sharding.chop_domain((4, 0, 4), 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, shape, mgs, out_dir):
    import torch.distributed as dist
    from oracle.pyoracle import BoxState, Oracle, reacteval_states
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = Oracle("LiDryer")
    st = reacteval_states(orc, shape)
    nx, ny, nz = shape
    N = st.rY.shape[0]
    rY = st.rY.reshape(N, nz, ny, nx)
    T, rE = st.T.reshape(nz, ny, nx), st.rE.reshape(nz, ny, nx)
    out = torch.zeros((N + 2, nz, ny, nx), dtype=torch.float64)
    boxes = sharding.chop_domain(shape, mgs)
    local = {k: 0 for k in sharding.SUM_KEYS + sharding.MAX_KEYS}
    for ib in sharding.my_boxes(boxes, rank, world):
        (x0, y0, z0), (x1, y1, z1) = boxes[ib]
        sl = (slice(z0, z1 + 1), slice(y0, y1 + 1), slice(x0, x1 + 1))
        nb = sharding.box_cells(boxes[ib])
        b = BoxState(rY[(slice(None),) + sl].reshape(N, nb).copy(), T[sl].reshape(nb).copy(), rE[sl].reshape(nb).copy(),
                     np.zeros((N, nb)), np.zeros(nb))
        orc.react_rk64(b, 1e-6, atol=1e-10)
        out[(slice(0, N),) + sl] = torch.from_numpy(b.rY.reshape((N,) + T[sl].shape))
        out[(N,) + sl] = torch.from_numpy(b.T.reshape(T[sl].shape))
        out[(N + 1,) + sl] = torch.from_numpy(b.FC.reshape(T[sl].shape))
        local["cells"] += nb
        local["n_steps"] += int(b.FC.sum())
        local["n_rhs"] += 6 * int(b.FC.sum())
        local["max_steps"] = max(local["max_steps"], int(b.FC.max()))
        local["device_ms"] = max(local["device_ms"], float(rank + 1))
    dist.all_reduce(out)  # test-only gather: every cell is written by exactly one rank
    red = sharding.reduce_stats(local)
    if rank == 0:
        np.save(os.path.join(out_dir, "out.npy"), out.numpy())
        np.save(os.path.join(out_dir, "stats.npy"), np.array([red[k] for k in ("cells", "n_steps", "n_rhs", "max_steps", "device_ms")]))
    dist.destroy_process_group()


def test_two_rank_sharding_reproduces_unsharded_result(tmp_path):
    import torch.multiprocessing as mp
    from oracle.pyoracle import Oracle, reacteval_states
    shape, mgs, world = (4, 8, 2), 4, 2
    mp.spawn(_worker, args=(world, _free_port(), shape, mgs, str(tmp_path)), nprocs=world, join=True)
    out = np.load(tmp_path / "out.npy")
    stats = np.load(tmp_path / "stats.npy")
    orc = Oracle("LiDryer")
    ref = reacteval_states(orc, shape)
    orc.react_rk64(ref, 1e-6, atol=1e-10)
    N = ref.rY.shape[0]
    assert np.array_equal(out[:N].reshape(N, -1), ref.rY)
    assert np.array_equal(out[N].ravel(), ref.T) and np.array_equal(out[N + 1].ravel(), ref.FC)
    assert stats[0] == ref.ncells and stats[1] == ref.FC.sum() and stats[2] == 6 * ref.FC.sum()
    assert stats[3] == ref.FC.max() and stats[4] == 2.0
