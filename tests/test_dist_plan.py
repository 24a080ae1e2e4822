"""Host-side logic of the row-partitioned path (liblcg_b200/dist.py) on CPU: world_size 2 and 3 over gloo.

Each rank remaps its rows, agrees on the halo plan with its peers, and the exchange the C library performs with
ncclSend/ncclRecv on the GPU is emulated here with the same plan over gloo: the partitioned SpMV and the
all-reduced dot must equal the global ones.  (The CUDA side of the same path is covered by tests/multi_gpu_check.py,
run under torchrun on a multi-GPU box.)
"""
import os
import socket

import numpy as np
import pytest

from liblcg_b200 import stencil


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _ragged(n, seed):
    rng = np.random.default_rng(seed)
    lens = rng.integers(0, 9, size=n)
    rp = np.zeros(n + 1, dtype=np.int32)
    np.cumsum(lens, out=rp[1:])
    col = np.concatenate([np.sort(rng.choice(n, size=k, replace=False)) for k in lens] + [np.zeros(0, dtype=np.int64)]).astype(np.int32)
    val = rng.standard_normal(len(col))
    return rp, col, val


def _worker(rank, world, port, case, out):
    import torch
    import torch.distributed as dist
    from liblcg_b200 import dist as ldist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        if case == "7pt":
            rp, col, val = stencil.make_stencil("7pt", 6)
            n = 6 ** 3
            bounds = ldist.row_bounds(n, world, align=36)
        elif case == "27pt":
            rp, col, val = stencil.make_stencil("27pt", 5)
            n = 5 ** 3
            bounds = ldist.row_bounds(n, world, align=25)
        else:
            n = 157
            rp, col, val = _ragged(n, 3)
            bounds = ldist.row_bounds(n, world)
        r0, r1 = bounds[rank], bounds[rank + 1]
        k0, k1 = rp[r0], rp[r1]
        new_col, plan = ldist.plan_partition(torch.from_numpy(col[k0:k1].astype(np.int64)), bounds, rank)
        new_col = new_col.numpy()
        n_loc = r1 - r0
        assert plan.n_local == n_loc and new_col.min(initial=0) >= 0 and new_col.max(initial=0) < n_loc + plan.n_ghost
        # ghosts are grouped by owner in ascending rank, ascending global id inside an owner
        gg = plan.ghost_global.numpy()
        assert np.all(np.diff(gg) > 0)
        assert sum(plan.recv_from.values()) == plan.n_ghost
        for p, idx in plan.send_to.items():
            assert p != rank and idx.min() >= 0 and idx.max() < n_loc
        # emulate the halo exchange with the plan
        x = np.random.default_rng(7).standard_normal(n)
        x_ext = np.zeros(n_loc + plan.n_ghost)
        x_ext[:n_loc] = x[r0:r1]
        outbox = {p: x_ext[idx] for p, idx in plan.send_to.items()}
        boxes = [None] * world
        dist.all_gather_object(boxes, outbox)
        off = n_loc
        for p in plan.peers:
            cnt = plan.recv_from.get(p, 0)
            if cnt:
                x_ext[off:off + cnt] = boxes[p][rank]
                off += cnt
        assert off == n_loc + plan.n_ghost
        np.testing.assert_array_equal(x_ext[n_loc:], x[gg])       # every ghost slot received the right entry
        # partitioned SpMV + all-reduced dot == global
        rpl = rp[r0:r1 + 1] - k0
        y_loc = np.array([np.dot(val[k0:k1][rpl[i]:rpl[i + 1]], x_ext[new_col[rpl[i]:rpl[i + 1]]]) for i in range(n_loc)])
        y_glob = np.array([np.dot(val[rp[i]:rp[i + 1]], x[col[rp[i]:rp[i + 1]]]) for i in range(n)])
        np.testing.assert_allclose(y_loc, y_glob[r0:r1], rtol=0, atol=1e-13)
        t = torch.tensor([float(np.dot(x[r0:r1], y_loc))], dtype=torch.float64)
        dist.all_reduce(t)
        assert abs(t.item() - float(np.dot(x, y_glob))) < 1e-10
        # rows [r0, r1) of A^T assembled from the row blocks of A (complex BiCG's A^H d2 on a partitioned system)
        import scipy.sparse as sp
        cval = val * (1.0 + 0.25j)
        t_rp, t_col, t_val = ldist.transposed_rows(rpl.astype(np.int32), col[k0:k1], cval[k0:k1], bounds, rank)
        At = sp.csr_matrix((cval, col, rp), shape=(n, n)).T.tocsr()
        At.sort_indices()
        assert np.array_equal(t_rp, At.indptr[r0:r1 + 1] - At.indptr[r0])
        assert np.array_equal(t_col, At.indices[At.indptr[r0]:At.indptr[r1]])
        np.testing.assert_array_equal(t_val, At.data[At.indptr[r0]:At.indptr[r1]])
        out[rank] = "ok"
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("case", ["7pt", "27pt", "ragged"])
def test_halo_plan_over_gloo(world, case):
    import torch.multiprocessing as mp
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, case, out), nprocs=world, join=True)
    assert [out.get(r) for r in range(world)] == ["ok"] * world


def test_row_bounds():
    from liblcg_b200 import dist as ldist
    assert ldist.row_bounds(10, 1) == [0, 10]
    assert ldist.row_bounds(8 ** 3, 8, align=64) == [64 * i for i in range(9)]
    b = ldist.row_bounds(256 ** 3, 3, align=256 * 256)
    assert b[0] == 0 and b[-1] == 256 ** 3 and all(x % 65536 == 0 for x in b) and all(b[i] < b[i + 1] for i in range(3))


def test_row_bounds_balanced_by_streamed_bytes():
    """contiguous row blocks balanced by nnz (bytes streamed per iteration), not by row count."""
    from liblcg_b200 import dist as ldist, stencil
    # a matrix whose second half is eight times denser: the cut moves towards the dense half's start
    lens = np.concatenate([np.full(500, 2), np.full(500, 16)])
    rp = np.zeros(1001, dtype=np.int64)
    np.cumsum(lens, out=rp[1:])
    b = ldist.row_bounds_nnz(rp, 2)
    w = ldist.NNZ_BYTES * rp + ldist.ROW_OVERHEAD_BYTES * np.arange(1001)
    assert 500 < b[1] < 1000 and abs(2 * w[b[1]] - w[-1]) <= 2 * (ldist.NNZ_BYTES * 16 + ldist.ROW_OVERHEAD_BYTES)
    assert ldist.row_bounds_nnz(rp, 1) == [0, 1000]
    b4 = ldist.row_bounds_nnz(rp, 4, align=10)
    assert b4[0] == 0 and b4[-1] == 1000 and all(x % 10 == 0 for x in b4) and all(b4[i] <= b4[i + 1] for i in range(4))
    # the analytic plane counts of the stencils are the generated matrices' row pointers at the plane edges
    for kind in ("7pt", "27pt", "7pt_cd"):
        for g in (5, 8):
            S = stencil.make_system(kind, g)
            assert np.array_equal(ldist.stencil_row_ptr_planes(kind, g), S["row_ptr"].astype(np.int64)[::g * g])
    # whole planes per rank, every rank non-empty, within one plane of the byte-balanced cut
    for world in (2, 3, 5, 8):
        bs = ldist.stencil_bounds("27pt", 32, world)
        assert bs[0] == 0 and bs[-1] == 32 ** 3 and all(x % 1024 == 0 for x in bs) and all(bs[i] < bs[i + 1] for i in range(world))
