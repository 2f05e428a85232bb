"""Host twin of the index logic of pp::ppext_tiled_kernel (cubep3m_b200/csrc/pp.cuh), checked against brute-force neighbour sets.

The kernel re-sorts the particles of a target block's source region (block of TB_X x TB_Y x TB_Z coarse cells + TB_HALO fine cells) by the
dense region cell index, derives the targets from a per-row table and walks, per target, a bit mask of non-empty neighbour rows. This test
replays exactly those steps in numpy — the coarse-row ranges of the sorted array, the region histogram / scan / fill, the 64-row target
table with its binary search, the row mask and the one-pair-per-iteration walk with the own-cell skip — and demands that every physical
particle gets exactly the reference's PP_EXT partner set (particle_mesh_threaded.f90:496-523: all particles of the cells within pp_range
in every dimension, the own cell excluded), once, for pp_range 1 and 2, including blocks clipped at the upper edge of the physical volume.
It pins the arithmetic that cannot be unit-tested on the device; the force values themselves are covered by tests/test_gpu_parity.py."""
import numpy as np
import pytest

TBX, TBY, TBZ, HALO = 4, 2, 2, 2                       # pp::TB_X, TB_Y, TB_Z, TB_HALO
RX, RY, RZ = 4 * TBX + 2 * HALO, 4 * TBY + 2 * HALO, 4 * TBZ + 2 * HALO
NCELL = RX * RY * RZ


def _sorted_particles(rng, n, nc_buf, nc_node):
    H, b, mT = nc_node + 2 * nc_buf, 4 * nc_buf, 4 * nc_node
    x = rng.uniform(-b, mT + b, size=(n, 3)).astype(np.float32)
    g = np.floor(x).astype(int) + b
    g = np.clip(g, 0, 4 * H - 1)
    key = (((g[:, 2] >> 2) * H + (g[:, 1] >> 2)) * H + (g[:, 0] >> 2)) * 64 + (((g[:, 2] & 3) << 4) | ((g[:, 1] & 3) << 2) | (g[:, 0] & 3))
    o = np.argsort(key, kind="stable")
    g, key = g[o], key[o]
    fstart = np.searchsorted(key, np.arange(H ** 3 * 64 + 1))
    return g, fstart, H


def _tiled_partner_sets(g, fstart, H, nc_buf, nc_node, pr):
    plo, phi = 4 * nc_buf, 4 * (nc_buf + nc_node)
    nbx, nby, nbz = -(-nc_node // TBX), -(-nc_node // TBY), -(-nc_node // TBZ)
    got = {}
    for blk in range(nbx * nby * nbz):
        bx, by, bz = blk % nbx, (blk // nbx) % nby, blk // (nbx * nby)
        cx0, cy0, cz0 = nc_buf + bx * TBX, nc_buf + by * TBY, nc_buf + bz * TBZ
        org = np.array([4 * cx0 - HALO, 4 * cy0 - HALO, 4 * cz0 - HALO])
        # 1./3. the (TB_Y+2) x (TB_Z+2) coarse x-rows, each one contiguous range of the sorted array
        recs = []
        for r in range((TBY + 2) * (TBZ + 2)):
            cy, cz = cy0 - 1 + r % (TBY + 2), cz0 - 1 + r // (TBY + 2)
            if cy > H - 1 or cz > H - 1:
                continue
            rk = (cz * H + cy) * H
            recs += range(fstart[(rk + cx0 - 1) * 64], fstart[(rk + min(cx0 + TBX, H - 1)) * 64 + 64])
        tab = np.zeros(NCELL + 1, int)
        cells = []
        for gi in recs:
            l = g[gi] - org
            if (l >= 0).all() and l[0] < RX and l[1] < RY and l[2] < RZ:
                c = (l[2] * RY + l[1]) * RX + l[0]
                cells.append((gi, c)); tab[c + 1] += 1
        # 2. exclusive scan in place on tab[1..]; the fill turns tab[c + 1] into the start of cell c + 1
        run = 0
        for c in range(1, NCELL + 1):
            v = tab[c]; tab[c] = run; run += v
        src = [None] * run
        for gi, c in cells:
            src[tab[c + 1]] = gi; tab[c + 1] += 1
        # 4. target rows: one contiguous range per interior (z,y) row
        xl, xh = max(HALO, plo - org[0]), min(HALO + 4 * TBX, phi - org[0])
        pre, st = [0], []
        for row in range(16 * TBY * TBZ):
            ly, lz = HALO + row % (4 * TBY), HALO + row // (4 * TBY)
            live = xh > xl and plo <= ly + org[1] < phi and plo <= lz + org[2] < phi
            rb = (lz * RY + ly) * RX
            s0 = tab[rb + xl] if live else 0
            st.append(s0); pre.append(pre[-1] + ((tab[rb + xh] - s0) if live else 0))
        # 5. walk
        w, qc = 2 * pr + 1, ((2 * pr + 1) ** 2 - 1) >> 1
        for t in range(pre[-1]):
            r, h = 0, 32
            while h >= 1:
                if pre[r + h] <= t:
                    r += h
                h //= 2
            gi = src[st[r] + (t - pre[r])]
            lx, ly, lz = g[gi] - org
            own = (lz * RY + ly) * RX + lx
            own_s, own_e = tab[own], tab[own + 1]
            rows, rb, q = 0, own - pr * (RY + 1) * RX, 0
            for qz in range(w):
                for qy in range(w):
                    if tab[rb + pr + 1] - tab[rb - pr] - ((own_e - own_s) if q == qc else 0) > 0:
                        rows |= 1 << q
                    q += 1; rb += RX
                rb += (RY - w) * RX
            acc, s, e = [], 0, 0
            while True:
                if s == own_s:
                    s = own_e
                if s >= e:
                    if not rows:
                        break
                    q = (rows & -rows).bit_length() - 1
                    rows &= rows - 1
                    qz = (q * 52) >> 8 if w == 5 else ((q * 86) >> 8 if w == 3 else 0)
                    qy = q - qz * w
                    rb = own + ((qz - pr) * RY + (qy - pr)) * RX
                    s, e = tab[rb - pr], tab[rb + pr + 1]
                    if s == own_s:
                        s = own_e
                assert s < e
                acc.append(src[s]); s += 1
            assert gi not in got, "a target was walked twice"
            assert len(set(acc)) == len(acc), "a partner was visited twice"
            got[gi] = set(acc)
    return got


@pytest.mark.parametrize("nc_node,pr,n", [(8, 2, 30000), (6, 2, 20000), (6, 1, 20000), (8, 1, 4000)])
def test_tiled_walk_visits_exactly_the_reference_partner_sets(nc_node, pr, n):
    nc_buf = 6
    g, fstart, H = _sorted_particles(np.random.default_rng(nc_node * 10 + pr), n, nc_buf, nc_node)
    got = _tiled_partner_sets(g, fstart, H, nc_buf, nc_node, pr)
    plo, phi = 4 * nc_buf, 4 * (nc_buf + nc_node)
    phys = np.nonzero(np.all((g >= plo) & (g < phi), axis=1))[0]
    assert set(got) == set(phys.tolist())
    for i in phys:
        d = np.abs(g - g[i])
        ref = set(np.nonzero(np.all(d <= pr, axis=1) & ~np.all(d == 0, axis=1))[0].tolist())
        assert got[i] == ref
