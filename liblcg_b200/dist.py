"""Row-partitioned systems over several GPUs: one process per GPU, `torch.distributed` for the
plumbing (rendezvous, exchanging the halo plan and the NCCL id), the C ABI (lcgb200_comm_*, lcgb200_csr_set_partition)
for everything on the iteration path.

The reference is single-device; this module is the out-of-band selection of the multi-GPU path the survey
describes: callers still drive the solvers through `api.solve` with their row slices of m and B.

Partition: contiguous row blocks [bounds[r], bounds[r+1]).  Each rank keeps its CSR rows with columns remapped to the
extended vector  [ local entries | ghost entries grouped by owning rank, ascending global index within an owner ].
The plan (who needs which of my rows) is derived from the column indices alone, so it works for any CSR, not only
the stencils.  `plan_partition` is pure tensor code (CPU or CUDA) and is what the gloo tests exercise.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

from . import _lib, api


def row_bounds(n: int, world: int, align: int = 1) -> list[int]:
    """Balanced contiguous row blocks; block edges rounded to a multiple of `align` (e.g. a grid plane)."""
    b = [0]
    for r in range(1, world):
        e = (n * r) // world
        if align > 1:
            e = int(round(e / align)) * align
        b.append(min(max(e, b[-1]), n))
    b.append(n)
    return b


# bytes one row costs per iteration besides its non-zeros: row_ptr, x, y of the SpMV (20 B) + ~11 vector passes of 8 B
ROW_OVERHEAD_BYTES = 108
NNZ_BYTES = 12


def row_bounds_nnz(row_ptr, world: int, align: int = 1) -> list[int]:
    """Contiguous row blocks balanced by the BYTES an iteration streams ("balanced by nnz"): a row weighs
    12 B per non-zero plus ROW_OVERHEAD_BYTES.  `row_ptr` is the global row pointer (numpy or tensor, n + 1 entries);
    block edges are rounded to a multiple of `align` rows (e.g. a grid plane, so that send lists stay contiguous)."""
    rp = np.asarray(row_ptr.cpu() if hasattr(row_ptr, "cpu") else row_ptr, dtype=np.int64)
    n = len(rp) - 1
    weight = NNZ_BYTES * (rp - rp[0]) + ROW_OVERHEAD_BYTES * np.arange(n + 1, dtype=np.int64)   # prefix sums of the row weights
    b = [0]
    for r in range(1, world):
        e = int(np.searchsorted(weight, weight[-1] * r / world, side="left"))
        if align > 1:
            e = int(round(e / align)) * align
        b.append(min(max(e, b[-1]), n))
    b.append(n)
    return b


def stencil_row_ptr_planes(kind: str, g: int) -> np.ndarray:
    """Prefix sums of the non-zeros per z-plane of the g^3 stencils (plane z holds rows [z g^2, (z+1) g^2)), as a row pointer
    over PLANES — what row_bounds_nnz needs to balance a z-slab partition without building the matrix."""
    def line(reach):   # neighbours within `reach` along one axis, summed over a line of g points
        return g + 2 * (g - 1) * reach
    if kind == "27pt":
        per_axis = np.array([3 if 0 < z < g - 1 else (2 if g > 1 else 1) for z in range(g)], dtype=np.int64)
        plane = per_axis * (line(1) ** 2)
    else:   # 7-point stars: centre + 2 in-plane axes + the z neighbours
        in_plane = g * g + 2 * 2 * g * (g - 1)
        plane = np.array([in_plane + g * g * ((z > 0) + (z < g - 1)) for z in range(g)], dtype=np.int64)
    out = np.zeros(g + 1, dtype=np.int64)
    np.cumsum(plane, out=out[1:])
    return out


def stencil_bounds(kind: str, g: int, world: int) -> list[int]:
    """z-slab partition of a g^3 stencil system balanced by streamed bytes (whole planes per rank)."""
    planes = stencil_row_ptr_planes(kind, g)
    weight = NNZ_BYTES * planes + ROW_OVERHEAD_BYTES * g * g * np.arange(g + 1, dtype=np.int64)
    b = [0]
    for r in range(1, world):
        z = int(np.searchsorted(weight, weight[-1] * r / world, side="left"))
        if z > 0 and abs(weight[z - 1] - weight[-1] * r / world) <= abs(weight[z] - weight[-1] * r / world):
            z -= 1
        b.append(min(max(z * g * g, b[-1]), g ** 3))
    b.append(g ** 3)
    return b


@dataclass
class HaloPlan:
    rank: int
    world: int
    r0: int
    r1: int
    n_global: int
    ghost_global: "object"            # sorted global column ids of my ghost entries (tensor)
    recv_from: dict = field(default_factory=dict)   # peer -> number of ghost entries it owns
    send_to: dict = field(default_factory=dict)     # peer -> np.int32 array of MY local row indices it needs

    @property
    def n_local(self):
        return self.r1 - self.r0

    @property
    def n_ghost(self):
        return int(self.ghost_global.numel())

    @property
    def peers(self):
        return sorted(set(self.recv_from) | set(self.send_to))


def remap_columns(col_global, r0: int, r1: int, bounds):
    """Local column indices for rows [r0,r1): (new_col int32, ghost_global sorted int64, ghost owner counts per rank)."""
    import torch
    col = col_global.to(torch.int64)
    outside = (col < r0) | (col >= r1)
    ghost = torch.unique(col[outside])                      # sorted
    new_col = col - r0
    if ghost.numel():
        pos = torch.searchsorted(ghost, col[outside])
        new_col[outside] = (r1 - r0) + pos
    edges = torch.tensor(bounds[1:], dtype=torch.int64, device=col.device)
    owner = torch.bucketize(ghost, edges, right=True)       # rank owning each ghost entry
    counts = torch.bincount(owner, minlength=len(bounds) - 1).cpu().tolist() if ghost.numel() else [0] * (len(bounds) - 1)
    return new_col.to(torch.int32), ghost, counts


def plan_partition(col_global, bounds, rank: int, group=None):
    """Remap the columns of this rank's rows and agree with the peers on the halo plan.
    Returns (new_col int32 tensor, HaloPlan).  Collective over `group` (any backend: only all_gather_object)."""
    import torch.distributed as dist
    world = len(bounds) - 1
    r0, r1 = bounds[rank], bounds[rank + 1]
    new_col, ghost, counts = remap_columns(col_global, r0, r1, bounds)
    plan = HaloPlan(rank, world, r0, r1, bounds[-1], ghost)
    need = {}
    off = 0
    g_cpu = ghost.cpu().numpy()
    for p in range(world):
        if counts[p]:
            need[p] = g_cpu[off:off + counts[p]]
            plan.recv_from[p] = counts[p]
            off += counts[p]
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, need, group=group)
        for p, needs_of_p in enumerate(gathered):
            if p != rank and needs_of_p and rank in needs_of_p:
                plan.send_to[p] = (np.asarray(needs_of_p[rank], dtype=np.int64) - r0).astype(np.int32)
    return new_col, plan


class Communicator:
    """lcgb200_comm_t: NCCL communicator owned by the C library; the id travels over torch.distributed."""

    def __init__(self, rank: int, world: int, group=None):
        import torch.distributed as dist
        lib = _lib.load()
        ident = [None]
        if rank == 0:
            buf = (C.c_ubyte * 128)()
            rc = lib.lcgb200_comm_unique_id(buf, 128)
            if rc != 0:
                raise RuntimeError(f"lcgb200_comm_unique_id failed ({rc}): {api.last_error()}")
            ident[0] = bytes(buf)
        dist.broadcast_object_list(ident, src=0, group=group)
        h = C.c_void_p()
        idbuf = (C.c_ubyte * 128).from_buffer_copy(ident[0])
        rc = lib.lcgb200_comm_create(C.byref(h), rank, world, idbuf)
        if rc != 0:
            raise RuntimeError(f"lcgb200_comm_create failed ({rc}): {api.last_error()}")
        self.handle, self.rank, self.world = h, rank, world

    def stats(self):
        a, b = C.c_int(), C.c_int()
        _lib.load().lcgb200_comm_stats(self.handle, C.byref(a), C.byref(b))
        return {"halo_exchanges": a.value, "allreduces": b.value}

    def close(self):
        if getattr(self, "handle", None):
            _lib.load().lcgb200_comm_destroy(self.handle)
            self.handle = None


@dataclass
class Partition:
    op: api.CsrOperator
    comm: Communicator
    plan: HaloPlan
    n_local: int
    b: "object" = None
    p2p: bool = False
    t_part: "object" = None     # the rows of A^T as a second partition (complex BiCG's A^H d2), attached to `op`

    def close(self):
        self.op.close()
        self.comm.close()
        if self.t_part is not None:
            self.t_part.close()
            self.t_part = None


def attach_plan(op: api.CsrOperator, comm: Communicator, plan: HaloPlan) -> None:
    peers = plan.peers
    ranks = np.asarray(peers, dtype=np.int32)
    send_counts = np.asarray([len(plan.send_to.get(p, ())) for p in peers], dtype=np.int32)
    recv_counts = np.asarray([plan.recv_from.get(p, 0) for p in peers], dtype=np.int32)
    send_idx = np.concatenate([plan.send_to[p] for p in peers if p in plan.send_to]).astype(np.int32) if send_counts.sum() else np.zeros(1, np.int32)
    rc = _lib.load().lcgb200_csr_set_partition(op.handle, comm.handle, plan.n_global, len(peers), ranks.ctypes.data,
                                               send_counts.ctypes.data, send_idx.ctypes.data, recv_counts.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"lcgb200_csr_set_partition failed ({rc}): {api.last_error()}")


def enable_p2p(comm: Communicator, plan: HaloPlan, group=None) -> bool:
    """NVLink peer-memory transport: exchange the CUDA-IPC handles of the ranks' communication windows and the offsets at
    which each rank's entries start inside its neighbours' ghost regions, then map the peers' windows.  Returns False
    (and leaves the NCCL transport in place) when the devices cannot map each other's memory."""
    import torch.distributed as dist
    lib = _lib.load()
    buf = (C.c_ubyte * 64)()
    ng = C.c_longlong()
    rc = lib.lcgb200_comm_p2p_handle(comm.handle, buf, C.byref(ng))
    # where the entries a peer sends me start inside MY ghost region (peer order = ascending rank = ghost order)
    my_off, off = {}, 0
    for p in plan.peers:
        my_off[p] = off
        off += plan.recv_from.get(p, 0)
    info = {"ok": rc == 0, "handle": bytes(buf), "n_ghost": int(ng.value), "recv_off": my_off}
    gathered = [None] * comm.world
    dist.all_gather_object(gathered, info, group=group)
    if not all(g["ok"] for g in gathered):
        return False
    handles = b"".join(g["handle"] for g in gathered)
    n_ghost = np.asarray([g["n_ghost"] for g in gathered], dtype=np.int64)
    remote_off = np.asarray([gathered[p]["recv_off"].get(comm.rank, 0) for p in plan.peers] or [0], dtype=np.int64)
    hb = (C.c_ubyte * len(handles)).from_buffer_copy(handles)
    rc = lib.lcgb200_comm_p2p_attach(comm.handle, hb, n_ghost.ctypes.data, remote_off.ctypes.data)
    ok = [None] * comm.world
    dist.all_gather_object(ok, rc == 0, group=group)
    if not all(ok):   # e.g. no CUDA-IPC between two of the devices: every rank falls back to the NCCL transport
        lib.lcgb200_comm_p2p_detach(comm.handle)
        return False
    return True


def transposed_rows(row_ptr_local, col_global, val, bounds, rank: int, group=None):
    """Rows [r0, r1) of A^T from the row blocks of A: every rank buckets its entries (i, j, v) by the owner of column j and
    ships them there (set-up time, over all_gather_object: any backend).  Returns numpy (row_ptr rebased to 0, GLOBAL column
    ids ascending within a row, values NOT conjugated)."""
    import torch
    import torch.distributed as dist
    world = len(bounds) - 1
    r0, r1 = bounds[rank], bounds[rank + 1]
    to_np = lambda a: a.detach().cpu().numpy() if hasattr(a, "data_ptr") else np.asarray(a)
    rp, cj, v = to_np(row_ptr_local).astype(np.int64), to_np(col_global).astype(np.int64), to_np(val)
    ri = np.repeat(np.arange(r0, r1, dtype=np.int64), np.diff(rp))
    owner = np.searchsorted(np.asarray(bounds[1:], dtype=np.int64), cj, side="right")
    out = {p: (cj[owner == p], ri[owner == p], v[owner == p]) for p in range(world) if np.any(owner == p)}
    gathered = [out]
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, out, group=group)
    mine = [g[rank] for g in gathered if g and rank in g]
    trow = np.concatenate([m[0] for m in mine]) if mine else np.zeros(0, np.int64)
    tcol = np.concatenate([m[1] for m in mine]) if mine else np.zeros(0, np.int64)
    tval = np.concatenate([m[2] for m in mine]) if mine else np.zeros(0, v.dtype)
    order = np.lexsort((tcol, trow))
    trow, tcol, tval = trow[order], tcol[order], tval[order]
    t_rp = np.zeros(r1 - r0 + 1, dtype=np.int32)
    np.cumsum(np.bincount(trow - r0, minlength=r1 - r0), out=t_rp[1:])
    return t_rp, tcol.astype(np.int32), tval


def partition_csr(row_ptr_local, col_global, val, bounds, rank: int, jacobi=False, group=None, compress=False, transpose=False) -> Partition:
    """This rank's rows (row_ptr rebased to 0, GLOBAL column ids, values; torch CUDA tensors or numpy arrays)
    -> rectangular operator + communicator + halo plan.  transpose=True also builds the rows of A^T as a second partition
    (its own plan and windows) and attaches it, so that complex BiCG's A^H d2 (clcg.cpp:188) works on the partitioned system."""
    import torch
    world = len(bounds) - 1
    on_dev = hasattr(col_global, "data_ptr")
    colt = col_global if on_dev else torch.from_numpy(np.ascontiguousarray(col_global))
    new_col, plan = plan_partition(colt, bounds, rank, group=group)
    n_loc = bounds[rank + 1] - bounds[rank]
    if on_dev:
        op = api.CsrOperator(row_ptr_local, new_col.contiguous(), val, n_cols=n_loc + plan.n_ghost, jacobi=jacobi, compress=compress)
    else:
        op = api.CsrOperator(np.asarray(row_ptr_local), new_col.numpy(), np.asarray(val), n_cols=n_loc + plan.n_ghost, jacobi=jacobi, compress=compress)
    comm = Communicator(rank, world, group=group)
    attach_plan(op, comm, plan)
    # the shadow residual of complex CGS/BICGSTAB/TFQMR is one rand() sequence over the whole vector: tell the block where it starts
    _lib.load().lcgb200_csr_set_row_offset(op.handle, bounds[rank])
    p2p = False
    if on_dev and os.environ.get("LCGB200_NO_P2P", "0") != "1":   # LCGB200_NO_P2P=1: NCCL transport (comparison runs)
        p2p = enable_p2p(comm, plan, group=group)
    part = Partition(op, comm, plan, n_loc)
    part.p2p = p2p
    if transpose:
        t_rp, t_col, t_val = transposed_rows(row_ptr_local, col_global, val, bounds, rank, group=group)
        if on_dev:
            dev = col_global.device
            t_rp, t_col, t_val = torch.from_numpy(t_rp).to(dev), torch.from_numpy(t_col).to(dev), torch.from_numpy(np.ascontiguousarray(t_val)).to(dev)
        part.t_part = partition_csr(t_rp, t_col, t_val, bounds, rank, group=group)
        rc = _lib.load().lcgb200_csr_attach_transpose(op.handle, part.t_part.op.handle)
        if rc != 0:
            raise RuntimeError(f"lcgb200_csr_attach_transpose failed ({rc}): {api.last_error()}")
    return part


KIND_ID = {"7pt": 0, "27pt": 1, "7pt_cd": 2}


def build_stencil_partition(kind: str, g: int, rank: int, world: int, device, jacobi=False, group=None, compress=False) -> Partition:
    """Rows of this rank of the g^3 stencil system (z-slab partition), generated on the device."""
    import torch
    lib = _lib.load()
    n = g ** 3
    bounds = stencil_bounds(kind, g, world)
    r0, r1 = bounds[rank], bounds[rank + 1]
    nz = C.c_longlong()
    assert lib.lcgb200_gen_stencil(KIND_ID[kind], g, r0, r1, None, None, None, 0, C.byref(nz), None) == 0
    rp = torch.empty(r1 - r0 + 1, dtype=torch.int32, device=device)
    ci = torch.empty(nz.value, dtype=torch.int32, device=device)
    va = torch.empty(nz.value, dtype=torch.float64, device=device)
    assert lib.lcgb200_gen_stencil(KIND_ID[kind], g, r0, r1, rp.data_ptr(), ci.data_ptr(), va.data_ptr(), 0, None, None) == 0
    b = torch.empty(r1 - r0, dtype=torch.float64, device=device)
    assert lib.lcgb200_gen_rhs(KIND_ID[kind], g, r0, r1, b.data_ptr(), None) == 0
    torch.cuda.synchronize()
    part = partition_csr(rp, ci, va, bounds, rank, jacobi=jacobi, group=group, compress=compress)
    part.b = b
    del rp, ci, va
    torch.cuda.empty_cache()
    return part
