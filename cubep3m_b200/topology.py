"""Rank-grid bookkeeping shared by bench.py and the tests (host logic only).

The reference decomposes into nodes_dim^3 cubic nodes with rank = x + D*y + D^2*z and periodic neighbours
(mpi_initialization.f90:42-76). Here the grid may be (Dx,Dy,Dz): 1 GPU (1,1,1); 2 GPUs (2,1,1); 4 GPUs (2,2,1); 8 GPUs (2,2,2) =
the reference's nodes_dim = 2. At 2 and 4 GPUs this is the north-star's "tile split": the tiles of one (non-cubic) box are
divided block-wise between the GPUs, each block being a cubic node of tiles_node_dim^3 tiles."""
import numpy as np

GRIDS = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}


def grid_for_world(n):
    if n not in GRIDS:
        raise ValueError(f"unsupported GPU count {n} (1, 2, 4 or 8)")
    return GRIDS[n]


def rank_coords(rank, grid):
    return (rank % grid[0], (rank // grid[0]) % grid[1], rank // (grid[0] * grid[1]))


def rank_of(coords, grid):
    x, y, z = (coords[i] % grid[i] for i in range(3))
    return x + grid[0] * (y + grid[1] * z)


def neighbours(rank, grid):
    """(-x, +x, -y, +y, -z, +z) ranks, periodic — mpi_cart_shift of mpi_initialization.f90:73-76."""
    c = list(rank_coords(rank, grid))
    out = []
    for a in range(3):
        for s in (-1, +1):
            d = c.copy()
            d[a] += s
            out.append(rank_of(d, grid))
    return tuple(out)


def split_global(xv_global, mT, grid):
    """Assign particles of a global (grid*mT)-sized box to ranks by position; returns local-coordinate lists per rank."""
    xv_global = np.asarray(xv_global, np.float32)
    c = np.floor(xv_global[:, :3] / np.float32(mT)).astype(np.int64)
    for a in range(3):
        c[:, a] = np.clip(c[:, a], 0, grid[a] - 1)
    r = c[:, 0] + grid[0] * (c[:, 1] + grid[1] * c[:, 2])
    out = []
    for k in range(grid[0] * grid[1] * grid[2]):
        p = xv_global[r == k].copy()
        cc = rank_coords(k, grid)
        for a in range(3):
            p[:, a] -= np.float32(cc[a] * mT)
        out.append(p)
    return out


def pass_axis_host(xv, axis, mT, nf_buf, eps, exchange):
    """Host restatement of one axis of particle_pass (particle_pass.f90:69-298) for one rank.
    `exchange(plus_going, minus_going) -> (from_minus, from_plus)` performs the neighbour exchange."""
    xv = np.asarray(xv, np.float32)
    fmT, b, e = np.float32(mT), np.float32(nf_buf), np.float32(eps)
    q = xv[:, axis]
    plus_going = xv[q >= fmT - b]
    minus_going = xv[q < b]
    from_minus, from_plus = exchange(plus_going, minus_going)
    a = from_minus.copy()
    if len(a):
        a[:, axis] = np.maximum(a[:, axis] - fmT, -b)                      # :162
    c = from_plus.copy()
    if len(c):
        v = c[:, axis]
        small = np.abs(v) < e
        v = np.where(small, np.where(v < 0, -e, e), v).astype(np.float32)   # :257-263
        c[:, axis] = np.minimum(v + fmT, (fmT + b) - e)                    # :264-265
    return np.concatenate([xv, a, c], axis=0)
