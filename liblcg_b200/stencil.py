"""Synthetic stencil systems exactly as SURVEY.md §8(d) defines them (host/numpy side).

Grid g^3, row i = (z*g + y)*g + x, neighbours outside the cube dropped (Dirichlet), columns ascending,
int32 indices, base 0.
  * '7pt'   Poisson:               diag 6,  six off-diagonals -1
  * '27pt'  Poisson:               diag 26, 26 off-diagonals -1
  * '7pt_cd' convection-diffusion: diag 6, off-diagonals -1-gamma_d (lower neighbour) / -1+gamma_d (upper
            neighbour), (gamma_x, gamma_y, gamma_z) = (0.5, 0.25, 0.125)  -> nonsymmetric
x*[i] = uint32(i * 2654435761) / 2^32, b = A x* accumulated in row order in double, m0 = 0.

The device-side generator (csrc/stencil_gen.cu, `lcgb200_gen_stencil`) produces the same arrays bit-for-bit
for a row range [row0, row1) and is what bench.py uses at 256^3 / 512^3; this file is the small-case
definition the tests compare it with.
"""
from __future__ import annotations

import numpy as np

KINDS = ("7pt", "27pt", "7pt_cd")
GAMMA = (0.5, 0.25, 0.125)


def stencil_nnz(kind: str, g: int) -> int:
    if kind in ("7pt", "7pt_cd"):
        return 7 * g**3 - 6 * g**2
    if kind == "27pt":
        return (3 * g - 2) ** 3
    raise ValueError(kind)


def _offsets(kind: str):
    """(dz, dy, dx, value) in ascending column order."""
    offs = []
    if kind == "27pt":
        for dz in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for dx in (-1, 0, 1):
                    offs.append((dz, dy, dx, 26.0 if (dz, dy, dx) == (0, 0, 0) else -1.0))
    else:
        cd = kind == "7pt_cd"
        gx, gy, gz = GAMMA if cd else (0.0, 0.0, 0.0)
        offs = [(-1, 0, 0, -1.0 - gz), (0, -1, 0, -1.0 - gy), (0, 0, -1, -1.0 - gx), (0, 0, 0, 6.0),
                (0, 0, 1, -1.0 + gx), (0, 1, 0, -1.0 + gy), (1, 0, 0, -1.0 + gz)]
    return offs


def make_stencil(kind: str, g: int, row0: int = 0, row1: int | None = None):
    """CSR rows [row0, row1) of the g^3 stencil matrix with GLOBAL column indices.
    Returns (row_ptr int32[rows+1], col int32[nnz], val float64[nnz])."""
    n = g**3
    row1 = n if row1 is None else row1
    rows = np.arange(row0, row1, dtype=np.int64)
    x = rows % g
    y = (rows // g) % g
    z = rows // (g * g)
    offs = _offsets(kind)
    k = len(offs)
    cols = np.empty((len(rows), k), dtype=np.int64)
    vals = np.empty((len(rows), k), dtype=np.float64)
    ok = np.empty((len(rows), k), dtype=bool)
    for j, (dz, dy, dx, v) in enumerate(offs):
        zz, yy, xx = z + dz, y + dy, x + dx
        ok[:, j] = (zz >= 0) & (zz < g) & (yy >= 0) & (yy < g) & (xx >= 0) & (xx < g)
        cols[:, j] = (zz * g + yy) * g + xx
        vals[:, j] = v
    counts = ok.sum(axis=1)
    row_ptr = np.zeros(len(rows) + 1, dtype=np.int64)
    np.cumsum(counts, out=row_ptr[1:])
    return row_ptr.astype(np.int32), cols[ok].astype(np.int32), vals[ok]


def x_star(n0: int, n1: int) -> np.ndarray:
    """x*[i] = uint32(i * 2654435761) / 2^32 for i in [n0, n1)."""
    i = np.arange(n0, n1, dtype=np.uint64)
    return ((i * np.uint64(2654435761)) & np.uint64(0xFFFFFFFF)).astype(np.float64) / 4294967296.0


def rhs_from_xstar(row_ptr, col, val) -> np.ndarray:
    """b = A x*, accumulated left to right within each row (serial order) in double."""
    n = len(row_ptr) - 1
    xs_all = x_star(0, int(col.max()) + 1 if len(col) else 0)
    b = np.zeros(n, dtype=np.float64)
    lens = np.diff(row_ptr)
    kmax = int(lens.max()) if n else 0
    for j in range(kmax):  # j-th entry of every row: keeps the per-row left-to-right order
        sel = lens > j
        idx = row_ptr[:-1][sel].astype(np.int64) + j
        b[sel] += val[idx] * xs_all[col[idx]]
    return b


def make_system(kind: str, g: int):
    """Full single-GPU system: dict(n, nnz, row_ptr, col, val, b, x_star)."""
    row_ptr, col, val = make_stencil(kind, g)
    n = g**3
    return dict(n=n, nnz=int(row_ptr[-1]), row_ptr=row_ptr, col=col, val=val,
                b=rhs_from_xstar(row_ptr, col, val), x_star=x_star(0, n))
