"""numpy restatement of the device generators (spmv_cache_trace_b200/csrc/generators.cu).

Test infrastructure: the GPU tests compare the device-generated matrices with these bit for
bit at sizes numpy handles in seconds.  Entries are returned 1-based (Matrix Market style),
row-major sorted, so they can be fed straight to the oracle's converters.
"""
import numpy as np

M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def stencil_entries(kind: int, nx: int, ny: int, nz: int = 1, row_begin: int = 0, row_end: int = 0):
    """kind 0: 2D 5-point (4,-1); 1: 3D 7-point (6,-1); 2: 3D 27-point (26,-1).  x fastest."""
    n = nx * ny * nz
    if row_end == 0 and row_begin == 0:
        row_end = n
    r = np.arange(row_begin, row_end, dtype=np.int64)
    ix, iy, iz = r % nx, (r // nx) % ny, r // (nx * ny)
    if kind == 2:
        offs = [(dz, dy, dx) for dz in (-1, 0, 1) for dy in (-1, 0, 1) for dx in (-1, 0, 1)]
        diag = 26.0
    elif kind == 1:
        offs = [(-1, 0, 0), (0, -1, 0), (0, 0, -1), (0, 0, 0), (0, 0, 1), (0, 1, 0), (1, 0, 0)]
        diag = 6.0
    else:
        offs = [(0, -1, 0), (0, 0, -1), (0, 0, 0), (0, 0, 1), (0, 1, 0)]
        diag = 4.0
    rows, cols, vals = [], [], []
    for dz, dy, dx in offs:
        ok = (ix + dx >= 0) & (ix + dx < nx) & (iy + dy >= 0) & (iy + dy < ny) & (iz + dz >= 0) & (iz + dz < nz)
        rows.append(r[ok])
        cols.append(r[ok] + (dz * ny + dy) * nx + dx)
        vals.append(np.full(int(ok.sum()), diag if (dx, dy, dz) == (0, 0, 0) else -1.0))
    rows, cols, vals = np.concatenate(rows), np.concatenate(cols), np.concatenate(vals)
    order = np.lexsort((cols, rows))
    rows, cols, vals = rows[order], cols[order], vals[order]
    return (rows - row_begin + 1).astype(np.int32), (cols + 1).astype(np.int32), vals


def splitmix64(z):
    z = np.asarray(z, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def rmat_entries(scale: int, edge_factor: int, seed: int, a=0.57, b=0.19, c=0.19):
    """Returns (rows, cols, vals) 0-based, row-major sorted, duplicates removed."""
    m = edge_factor << scale
    e = np.arange(m, dtype=np.uint64)
    ta, tb, tc = a, a + b, a + b + c
    row = np.zeros(m, dtype=np.uint64)
    col = np.zeros(m, dtype=np.uint64)
    with np.errstate(over="ignore"):
        for level in range(scale):
            h = splitmix64(np.uint64(seed) + e * np.uint64(0x9E3779B97F4A7C15)
                           + np.uint64(level) * np.uint64(0xBF58476D1CE4E5B9))
            u = (h >> np.uint64(11)).astype(np.float64) * 2.0 ** -53
            q = np.where(u < ta, 0, np.where(u < tb, 1, np.where(u < tc, 2, 3))).astype(np.uint64)
            row = (row << np.uint64(1)) | (q >> np.uint64(1))
            col = (col << np.uint64(1)) | (q & np.uint64(1))
    keys = np.unique((row << np.uint64(32)) | col)
    h = splitmix64(keys ^ np.uint64(0xD1B54A32D192ED03))
    vals = 2.0 * ((h >> np.uint64(11)).astype(np.float64) * 2.0 ** -53) - 1.0
    return (keys >> np.uint64(32)).astype(np.int64), (keys & np.uint64(0xFFFFFFFF)).astype(np.int64), vals
