"""Host twin of the slab-decomposed coarse-mesh solve (coarse_force.f90 + fft_coarse.f90 / fftw3ds.f90 on D^3 ranks).

NOT on the product path yet: the CUDA library solves the all-gathered global coarse mesh on every GPU. This module fixes the
communication plan of the replacement — which contiguous chunk goes to whom in each of its four exchanges — and executes it with numpy for
the local passes and a caller-supplied `exchange` for the messages, so that the plan is tested on the CPU (simulated ranks and a world-2
gloo run, tests/test_coarse_slab_plan.py) before the NCCL version is written against it.

Global coarse mesh (Nx,Ny,Nz) = nc * (Dx,Dy,Dz), W = Dx*Dy*Dz ranks, rank = x + Dx*(y + Dy*z) (cubep3m_b200/topology.py).
  1. cube -> z-slabs: rank q owns the zs = Nz/W planes [q*zs, (q+1)*zs). A cube rho_c[z][y][x] (z slowest) is nc/zs contiguous chunks of
     zs planes; chunk t of the cube at (rx,ry,rz) goes to rank rz*(nc/zs) + t, which places it at (x0,y0) = (rx*nc, ry*nc) of its slab.
  2. slab: r2c along x and FFT along y (local).
  3. transpose to y-pencils: rank s owns the ys = Ny/W rows [s*ys, (s+1)*ys) for ALL z. q sends slab[:, s*ys:(s+1)*ys, :] to s; s stores the
     block from q at planes [q*zs, (q+1)*zs) of T[z][yl][kx] — contiguous, no unpack.
  4. pencils: FFT along z, multiply by i*kern_c(comp) restricted to the rank's rows, inverse FFT along z (local), for the 3 components.
  5. transpose back: s sends G[q*zs:(q+1)*zs] (contiguous) to q, which stores it at rows [s*ys, (s+1)*ys) of its slab.
  6. slab: inverse FFT along y, c2r along x, / (Nx*Ny*Nz) (fftw3ds.f90:161).
  7. slab -> cube + 1-cell halo (what unpack_slab + coarse_force_buffer.f90:23-63 produce): rank q sends to every rank d the planes of its slab
     that fall in d's periodic z-range [rz*nc - 1, rz*nc + nc], each cut to d's periodic (nc+2) x (nc+2) window in y and x.
Requirements: zs = Nz/W divides nc; ys = Ny/W integer (true for nc a multiple of W on the supported grids)."""
import numpy as np

from . import topology as topo


class SlabPlan:
    def __init__(self, grid, nc):
        self.grid = tuple(int(g) for g in grid)
        self.nc = int(nc)
        self.W = self.grid[0] * self.grid[1] * self.grid[2]
        self.N = tuple(self.nc * g for g in self.grid)          # (Nx, Ny, Nz)
        if self.N[2] % self.W or self.N[1] % self.W:
            raise ValueError("Nz and Ny must be multiples of the rank count")
        self.zs, self.ys = self.N[2] // self.W, self.N[1] // self.W
        if self.nc % self.zs:
            raise ValueError("the slab thickness must divide nc_node")
        self.hc = self.N[0] // 2 + 1

    # ---- 1. cube -> slab
    def cube_sends(self, rank):
        """[(dest rank, first local z plane of the chunk)] — each chunk is rho_c[z0:z0+zs], contiguous."""
        rz = topo.rank_coords(rank, self.grid)[2]
        per = self.nc // self.zs
        return [(rz * per + t, t * self.zs) for t in range(per)]

    def cube_sources(self, rank):
        """[(source rank, x0, y0)] of the chunks rank receives for its slab."""
        rz = (rank * self.zs) // self.nc
        return [(topo.rank_of((rx, ry, rz), self.grid), rx * self.nc, ry * self.nc)
                for ry in range(self.grid[1]) for rx in range(self.grid[0])]

    # ---- 7. slab -> cube + halo
    def halo_planes(self, owner, dest):
        """[(slab-local plane, halo-local z index 0..nc+1)] of owner's slab that rank dest needs (periodic)."""
        rz = topo.rank_coords(dest, self.grid)[2]
        out = []
        for hz in range(self.nc + 2):
            z = (rz * self.nc - 1 + hz) % self.N[2]
            if z // self.zs == owner:
                out.append((z - owner * self.zs, hz))
        return out

    def halo_window(self, dest):
        """periodic global (y indices, x indices) of dest's (nc+2) x (nc+2) window."""
        rx, ry, _ = topo.rank_coords(dest, self.grid)
        return ((ry * self.nc - 1 + np.arange(self.nc + 2)) % self.N[1], (rx * self.nc - 1 + np.arange(self.nc + 2)) % self.N[0])


def solve_rank(plan, rank, rho_c, kern_rows, exchange):
    """One rank's part of the coarse solve. rho_c: (nc,nc,nc) [z][y][x]; kern_rows: (3, Nz, ys, hc) = kern_c[comp][z][rank's rows][kx];
    exchange(tag, {dest: array}) -> {source: array} delivers the messages of one exchange step. Returns force_c (3, nc+2, nc+2, nc+2)."""
    nc, zs, ys, hc, W = plan.nc, plan.zs, plan.ys, plan.hc, plan.W
    Nx, Ny, Nz = plan.N
    # 1. cube -> slab
    got = exchange("cube", {d: rho_c[z0:z0 + zs] for d, z0 in plan.cube_sends(rank)})
    slab = np.zeros((zs, Ny, Nx), np.float64)
    for src, x0, y0 in plan.cube_sources(rank):
        slab[:, y0:y0 + nc, x0:x0 + nc] = got[src]
    # 2. x r2c, y
    spec = np.fft.fft(np.fft.rfft(slab, axis=2), axis=1)                      # (zs, Ny, hc)
    # 3. transpose to pencils
    got = exchange("fwd", {s: spec[:, s * ys:(s + 1) * ys, :] for s in range(W)})
    T = np.concatenate([got[q] for q in range(W)], axis=0)                    # (Nz, ys, hc): block of q at planes q*zs..
    # 4. z forward, multiply, z inverse
    S = np.fft.fft(T, axis=0)
    force = np.zeros((3, nc + 2, nc + 2, nc + 2), np.float32)
    for comp in range(3):
        G = np.fft.ifft(1j * kern_rows[comp] * S, axis=0) * Nz                # unnormalised inverse along z
        # 5. transpose back
        got = exchange(f"bwd{comp}", {q: G[q * zs:(q + 1) * zs] for q in range(W)})
        back = np.concatenate([got[s] for s in range(W)], axis=1)             # (zs, Ny, hc): block of s at rows s*ys..
        # 6. y inverse, x c2r, normalisation
        real = np.fft.irfft(np.fft.ifft(back, axis=1) * Ny, n=Nx, axis=2) * Nx / (float(Nx) * Ny * Nz)
        # 7. slab -> cube + halo
        msgs = {}
        for d in range(W):
            planes = plan.halo_planes(rank, d)
            if planes:
                yi, xi = plan.halo_window(d)
                msgs[d] = real[[p for p, _ in planes]][:, yi][:, :, xi]
        got = exchange(f"halo{comp}", msgs)
        for owner, blk in got.items():
            for row, (_, hz) in enumerate(plan.halo_planes(owner, rank)):
                force[comp, hz] = blk[row]
    return force


def solve_simulated(plan, rho_cubes, kern_c):
    """All ranks in one process: runs solve_rank for every rank in lock-step (generators would hide the plan; a mailbox does not).
    kern_c: (3, Nz, Ny, hc) global. Returns [force_c per rank]."""
    import threading
    W = plan.W
    box, cond = {}, threading.Condition()
    out = [None] * W

    def make_exchange(rank):
        def exchange(tag, msgs):
            with cond:
                for d, a in msgs.items():
                    box[(tag, rank, d)] = np.array(a)
                box[(tag, "done", rank)] = True
                cond.notify_all()
                cond.wait_for(lambda: all((tag, "done", r) in box for r in range(W)))
                return {s: box[(tag, s, rank)] for s in range(W) if (tag, s, rank) in box}
        return exchange

    def run(rank):
        rows = kern_c[:, :, rank * plan.ys:(rank + 1) * plan.ys, :]
        out[rank] = solve_rank(plan, rank, rho_cubes[rank], rows, make_exchange(rank))

    th = [threading.Thread(target=run, args=(r,)) for r in range(W)]
    [t.start() for t in th]
    [t.join() for t in th]
    return out


def solve_global(plan, rho_cubes, kern_c):
    """Reference: assemble the global mesh, one global FFT solve, cut every rank's cube + periodic halo (what the library does today)."""
    nc = plan.nc
    Nx, Ny, Nz = plan.N
    rho = np.zeros((Nz, Ny, Nx), np.float64)
    for r, cube in enumerate(rho_cubes):
        rx, ry, rz = topo.rank_coords(r, plan.grid)
        rho[rz * nc:(rz + 1) * nc, ry * nc:(ry + 1) * nc, rx * nc:(rx + 1) * nc] = cube
    S = np.fft.fftn(np.fft.rfft(rho, axis=2), axes=(0, 1))
    out = []
    for r in range(plan.W):
        rx, ry, rz = topo.rank_coords(r, plan.grid)
        zi = (rz * nc - 1 + np.arange(nc + 2)) % Nz
        yi, xi = plan.halo_window(r)
        f = np.zeros((3, nc + 2, nc + 2, nc + 2), np.float32)
        for comp in range(3):
            real = np.fft.irfft(np.fft.ifftn(1j * kern_c[comp] * S, axes=(0, 1)), n=Nx, axis=2)
            f[comp] = real[zi][:, yi][:, :, xi]
        out.append(f)
    return out
