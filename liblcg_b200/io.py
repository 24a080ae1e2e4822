"""Front-of-path data step: liblcg's binary COO fixtures -> CSR.

File format (reference data/README:1-10, readers sample8.cu:30-64 and sample9.cu:30-64):
    case_*_A : int32 N | int32 nz | nz x { int32 row, int32 col, val } | N x b      (val/b: f64 or complex128)
    case_*_B : int32 N | N x x                                                      (the known answer)
The `[d]` block the README mentions is absent from the shipped files (sizes match without it).

The COO triplets are row-major sorted in the shipped files; `coo_to_csr` does not rely on it
(stable counting sort by row, the same result cusparseXcoo2csr gives on sorted input, sample8.cu:169).
"""
from __future__ import annotations

import os
import numpy as np

GOLDEN_DATA = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "data")


def coo_to_csr(n: int, rows: np.ndarray, cols: np.ndarray, vals: np.ndarray):
    """Stable COO -> CSR (int32 row_ptr[n+1], int32 col[nnz], val[nnz]); keeps duplicates, keeps in-row order."""
    rows = np.asarray(rows, dtype=np.int64)
    order = np.argsort(rows, kind="stable")
    counts = np.bincount(rows, minlength=n)
    row_ptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(counts, out=row_ptr[1:])
    if row_ptr[-1] >= 2**31:
        raise ValueError("nnz does not fit the 32-bit indices of the liblcg CUDA interface (lcg_cuda.h:81-83)")
    return row_ptr.astype(np.int32), np.ascontiguousarray(cols[order], dtype=np.int32), np.ascontiguousarray(vals[order])


def read_case(path_a: str, path_b: str | None = None, complex_valued: bool = False):
    """Read a reference fixture pair.  Returns dict(n, nnz, row_ptr, col, val, b, answer)."""
    vt = np.complex128 if complex_valued else np.float64
    raw = np.fromfile(path_a, dtype=np.uint8)
    n, nz = np.frombuffer(raw[:8].tobytes(), dtype=np.int32)
    n, nz = int(n), int(nz)
    rec = np.dtype([("r", "<i4"), ("c", "<i4"), ("v", vt)])
    off = 8
    trip = np.frombuffer(raw[off:off + nz * rec.itemsize].tobytes(), dtype=rec)
    off += nz * rec.itemsize
    b = np.frombuffer(raw[off:off + n * np.dtype(vt).itemsize].tobytes(), dtype=vt).copy()
    off += n * np.dtype(vt).itemsize
    if off != raw.size:
        raise ValueError(f"{path_a}: {raw.size - off} trailing bytes (unexpected layout)")
    row_ptr, col, val = coo_to_csr(n, trip["r"], trip["c"], trip["v"])
    out = dict(n=n, nnz=nz, row_ptr=row_ptr, col=col, val=val, b=b, answer=None)
    if path_b is not None:
        rawb = np.fromfile(path_b, dtype=np.uint8)
        nb = int(np.frombuffer(rawb[:4].tobytes(), dtype=np.int32)[0])
        if nb != n:
            raise ValueError("answer size does not match the system size")
        out["answer"] = np.frombuffer(rawb[4:4 + n * np.dtype(vt).itemsize].tobytes(), dtype=vt).copy()
    return out


def load_fixture(name: str):
    """name in {'10K', '10Kc', '1Kc'} -> the reference's data/case_* pairs committed under tests/golden/data."""
    table = {"10K": ("case_10K_A", "case_10K_B", False), "10Kc": ("case_10K_cA", "case_10K_cB", True),
             "1Kc": ("case_1K_cA", "case_1K_cB", True)}
    a, b, cx = table[name]
    return read_case(os.path.join(GOLDEN_DATA, a), os.path.join(GOLDEN_DATA, b), cx)


def csr_diagonal(row_ptr, col, val):
    """diag[i] = A[i,i] (first match per row, like lcg_smDcsr_get_diagonal_device, algebra_cuda.cu:40-57)."""
    n = len(row_ptr) - 1
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(row_ptr))
    hit = np.nonzero(col == rows)[0]
    diag = np.zeros(n, dtype=val.dtype)
    # keep the first hit per row
    r = rows[hit]
    first = np.ones(len(hit), dtype=bool)
    first[1:] = r[1:] != r[:-1]
    diag[r[first]] = val[hit[first]]
    return diag
