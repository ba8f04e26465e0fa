"""Deterministic synthetic maps / agent populations for the configs of BASELINE.json
(SURVEY.md 8d).  RNG: SplitMix64 -> xoshiro256**, seed 0x5EED0000 + config#.

Everything here is plain numpy: it feeds the CUDA path, the oracle and the reference alike."""
import numpy as np

MASK = (1 << 64) - 1
NAV_TILE = 4.0          # world units per nav tile (nav.c:4653-4661)
CHUNK_WU = 256.0        # world units per chunk


class Xoshiro:
    def __init__(self, seed):
        s = seed & MASK
        st = []
        for _ in range(4):                       # SplitMix64
            s = (s + 0x9E3779B97F4A7C15) & MASK
            z = s
            z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK
            z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK
            st.append(z ^ (z >> 31))
        self.s = st

    @staticmethod
    def _rotl(x, k):
        return ((x << k) | (x >> (64 - k))) & MASK

    def next(self):
        s = self.s
        r = (self._rotl((s[1] * 5) & MASK, 7) * 9) & MASK
        t = (s[1] << 17) & MASK
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]
        s[2] ^= t
        s[3] = self._rotl(s[3], 45)
        return r

    def randint(self, lo, hi):
        """uniform in [lo, hi)"""
        return lo + self.next() % (hi - lo)

    def uniform(self):
        return (self.next() >> 11) * (1.0 / (1 << 53))

    def numpy(self):
        """A numpy Generator seeded from this stream (bulk draws)."""
        return np.random.Generator(np.random.PCG64(self.next()))


def make_map(chunk_w, chunk_h, seed, frac_blocked=0.15, rivers=True, max_rect=12):
    """pathable[chunk_h*32][chunk_w*32] (map tiles; each becomes 2x2 nav tiles, nav.c:267-344).
    Seeded rectangles + straight rivers with fords; rectangle sizes in map tiles."""
    rng = Xoshiro(seed)
    H, W = chunk_h * 32, chunk_w * 32
    p = np.ones((H, W), dtype=np.uint8)
    target = int(frac_blocked * H * W)
    guard = 0
    while (p == 0).sum() < target and guard < 100000:
        guard += 1
        h = rng.randint(1, max_rect + 1); w = rng.randint(1, max_rect + 1)
        r = rng.randint(0, H); c = rng.randint(0, W)
        p[r:r + h, c:c + w] = 0
    if rivers and chunk_w * chunk_h > 1:
        nriv = max(1, (chunk_w + chunk_h) // 8)
        for i in range(nriv):
            if i % 2 == 0:
                r = rng.randint(H // 8, H - H // 8)
                p[r:r + 2, :] = 0
                for _ in range(max(2, chunk_w // 2)):
                    c = rng.randint(0, W - 6)
                    p[r:r + 2, c:c + 6] = 1          # ford
            else:
                c = rng.randint(W // 8, W - W // 8)
                p[:, c:c + 2] = 0
                for _ in range(max(2, chunk_h // 2)):
                    r = rng.randint(0, H - 6)
                    p[r:r + 6, c:c + 2] = 1
    return p


def cost_from_pathable(pathable, chunk_w, chunk_h):
    """cost_base in the chunk-blocked layout [chunk_r*chunk_w+chunk_c][64][64]: all tiles are FLAT
    with equal base_height, so cost = 1 where pathable else 0xFF (n_set_cost_for_tile, nav.c:267)."""
    nav = np.repeat(np.repeat(pathable, 2, axis=0), 2, axis=1)
    cost = np.where(nav != 0, 1, 0xFF).astype(np.uint8)
    return image_to_blocked(cost, chunk_w, chunk_h)


def image_to_blocked(img, chunk_w, chunk_h):
    return np.ascontiguousarray(
        img.reshape(chunk_h, 64, chunk_w, 64).transpose(0, 2, 1, 3).reshape(chunk_h * chunk_w, 64, 64))


def blocked_to_image(blk, chunk_w, chunk_h):
    return np.ascontiguousarray(
        blk.reshape(chunk_h, chunk_w, 64, 64).transpose(0, 2, 1, 3).reshape(chunk_h * 64, chunk_w * 64))


def tile_center_xz(chunk_r, chunk_c, tile_r, tile_c, map_x=0.0, map_z=0.0):
    """World xz of a nav tile centre (M_Tile_Bounds, tile.c:356: x decreases with the column)."""
    x = map_x - chunk_c * CHUNK_WU - tile_c * NAV_TILE - NAV_TILE / 2
    z = map_z + chunk_r * CHUNK_WU + tile_r * NAV_TILE + NAV_TILE / 2
    return np.float32(x), np.float32(z)


def tile_for_xz(x, z, chunk_w, chunk_h, map_x=0.0, map_z=0.0):
    gc = np.clip(((map_x - x) / NAV_TILE).astype(np.int64), 0, chunk_w * 64 - 1)
    gr = np.clip(((z - map_z) / NAV_TILE).astype(np.int64), 0, chunk_h * 64 - 1)
    return gr, gc


def random_passable_tiles(cost_blocked, n, rng):
    """n (chunk_idx, r, c) triples on passable tiles"""
    idx = np.argwhere(cost_blocked != 0xFF)
    sel = rng.integers(0, len(idx), size=n)
    return idx[sel]


def make_agents(cost_blocked, chunk_w, chunk_h, n, nflocks, seed, radius=1.0, max_speed=20.0,
                spacing=2.6, hz=20, map_x=0.0, map_z=0.0, goal_min_dist=150.0, cells=None):
    """Agent population: `nflocks` discs of agents on passable ground, hex-packed at `spacing` x radius,
    each flock with a seeded goal tile at least `goal_min_dist` wu away (so group arrival stays
    inactive, arrival.c:57).  Returns dict of numpy arrays.
    cells = (per_side, first): flock f is confined to cell `first + f` of a per_side x per_side grid over the map and
    spawns around that cell's centre, so flocks never overlap and the local density does not depend on how many
    flocks exist (weak-scaling populations)."""
    rng = Xoshiro(seed)
    g = rng.numpy()
    img = blocked_to_image(cost_blocked, chunk_w, chunk_h)
    H64, W64 = img.shape
    per = [n // nflocks + (1 if i < n % nflocks else 0) for i in range(nflocks)]
    pos = np.zeros((n, 2), np.float32)
    flock_of = np.zeros(n, np.int32)
    targets = np.zeros((nflocks, 2), np.float32)
    target_tiles = np.zeros((nflocks, 4), np.int32)
    passable = np.argwhere(img != 0xFF)
    k = 0
    radii = np.broadcast_to(np.asarray(radius, np.float32), (nflocks,)) if np.ndim(radius) <= 1 and np.size(radius) in (1, nflocks) \
        else np.asarray(radius, np.float32)
    rad_agent = np.zeros(n, np.float32)
    for f in range(nflocks):
        step = spacing * float(radii[f])
        # spawn centre
        cr, cc = passable[g.integers(0, len(passable))]
        bx0 = bz0 = -np.inf; bx1 = bz1 = np.inf
        if cells is not None:
            per_side, first = cells
            cell = first + f
            assert cell < per_side * per_side, "more flocks than grid cells"
            cell_w, cell_h = W64 * NAV_TILE / per_side, H64 * NAV_TILE / per_side
            gr_, gc_ = cell // per_side, cell % per_side
            bx1 = map_x - gc_ * cell_w; bx0 = bx1 - cell_w          # x decreases with the column
            bz0 = map_z + gr_ * cell_h; bz1 = bz0 + cell_h
            # nearest passable tile to the cell centre
            ctr_r = (gr_ + 0.5) * cell_h / NAV_TILE; ctr_c = (gc_ + 0.5) * cell_w / NAV_TILE
            k_ = np.argmin((passable[:, 0] - ctr_r) ** 2 + (passable[:, 1] - ctr_c) ** 2)
            cr, cc = passable[k_]
        cx = map_x - (cc + 0.5) * NAV_TILE
        cz = map_z + (cr + 0.5) * NAV_TILE
        need = per[f]
        got = 0
        shrinks = 0
        # hex spiral lattice around the centre, keeping only points on passable tiles inside the map
        R = int(np.ceil(np.sqrt(need / 0.55))) + 4
        while got < need:
            ii, jj = np.meshgrid(np.arange(-R, R + 1), np.arange(-R, R + 1), indexing="ij")
            xs = cx + (ii + 0.5 * (jj & 1)) * step
            zs = cz + jj * step * 0.8660254
            d2 = (xs - cx) ** 2 + (zs - cz) ** 2
            order = np.argsort(d2, axis=None, kind="stable")
            xs = xs.ravel()[order]; zs = zs.ravel()[order]
            inside = (xs < map_x - 1) & (xs > map_x - W64 * NAV_TILE + 1) & (zs > map_z + 1) & (zs < map_z + H64 * NAV_TILE - 1)
            inside &= (xs > bx0 + 1) & (xs < bx1 - 1) & (zs > bz0 + 1) & (zs < bz1 - 1)
            xs = xs[inside]; zs = zs[inside]
            tr, tc = tile_for_xz(xs, zs, chunk_w, chunk_h, map_x, map_z)
            ok = img[tr, tc] != 0xFF
            xs = xs[ok]; zs = zs[ok]
            got = len(xs)
            if got < need and R * step > (1.5 * max(bx1 - bx0, bz1 - bz0) if cells is not None else 4 * max(W64, H64) * NAV_TILE):
                # the whole area was searched. A flock confined to a grid cell whose passable part is too small for the
                # nominal spacing packs a little tighter (3 % per try) instead of failing: the weak-scaling populations
                # fill ~94 % of the passable map at N = 8, some cells hold more obstacles than others
                if cells is None or shrinks >= 40:
                    raise ValueError("flock %d: %d agents at spacing %.2f do not fit its area" % (f, need, step))
                shrinks += 1
                step *= 0.97
                R = int(np.ceil(np.sqrt(need / 0.55))) + 4
                continue
            R *= 2
        pos[k:k + need, 0] = xs[:need]
        pos[k:k + need, 1] = zs[:need]
        flock_of[k:k + need] = f
        rad_agent[k:k + need] = radii[f]
        # goal tile far enough from the spawn centre
        for _ in range(10000):
            tr_, tc_ = passable[g.integers(0, len(passable))]
            tx = map_x - (tc_ + 0.5) * NAV_TILE
            tz = map_z + (tr_ + 0.5) * NAV_TILE
            if (tx - cx) ** 2 + (tz - cz) ** 2 >= goal_min_dist ** 2 or chunk_w * chunk_h == 1:
                break
        targets[f] = (tx, tz)
        target_tiles[f] = (tr_ // 64, tc_ // 64, tr_ % 64, tc_ % 64)
        k += need
    # initial velocity: half max speed per tick towards the goal; prev_pos = pos - velocity
    tgt = targets[flock_of]
    d = tgt - pos
    ln = np.maximum(np.sqrt((d * d).sum(1, keepdims=True)), 1e-6)
    vel = (d / ln * (0.5 * max_speed / hz)).astype(np.float32)
    jitter = (g.random((n, 2)).astype(np.float32) - 0.5) * np.float32(0.05)
    vel = (vel + jitter).astype(np.float32)
    prev = (pos - vel).astype(np.float32)
    return dict(pos=pos, prev_pos=prev, vel=vel, radius=rad_agent,
                max_speed=np.full(n, max_speed, np.float32), speed=np.full(n, max_speed, np.float32),
                state=np.zeros(n, np.int32), flags=np.full(n, 1 << 3, np.uint32), flock_of=flock_of,
                flock_target=targets, flock_target_tile=target_tiles, hz=hz)
