"""Seeded scenario builders shared by the golden-vector generator, the CPU tests and the GPU tests.
Everything is a pure function of its seed (numpy + the package's synth module)."""
import importlib
import numpy as np

pf = importlib.import_module("permafrost-engine_b200")
capi, synth = pf.capi, pf.synth


def noise_map(cw, ch, seed, density):
    rng = np.random.default_rng(seed)
    p = synth.make_map(cw, ch, seed, frac_blocked=0.12, rivers=(cw * ch > 1))
    if density > 0:
        p[rng.random(p.shape) < density] = 0
    return p


def local_islands_np(cost, blockers=None):
    """numpy/python restatement of n_update_local_islands (nav.c:967) for test inputs: ids from 1 in
    row-major discovery order."""
    out = np.full(cost.shape, 0xFFFF, np.uint16)
    for ch in range(cost.shape[0]):
        nxt = 0
        free = cost[ch] != 0xFF
        if blockers is not None:
            free &= blockers[ch] == 0
        lab = out[ch]
        for r, c in np.argwhere(free):
            if lab[r, c] != 0xFFFF:
                continue
            nxt += 1
            stack = [(r, c)]
            lab[r, c] = nxt
            while stack:
                y, x = stack.pop()
                for yy, xx in ((y, x - 1), (y, x + 1), (y - 1, x), (y + 1, x)):
                    if 0 <= yy < 64 and 0 <= xx < 64 and free[yy, xx] and lab[yy, xx] == 0xFFFF:
                        lab[yy, xx] = nxt
                        stack.append((yy, xx))
    return out


def flow_tile_case(seed, dens, n=8):
    """-> cost[1,64,64], reqs"""
    p = noise_map(1, 1, seed, dens)
    cost = synth.cost_from_pathable(p, 1, 1)
    rng = np.random.default_rng(seed)
    tiles = np.argwhere(cost[0] != 255)
    sel = tiles[rng.integers(0, len(tiles), n)]
    if (cost[0] == 255).any():
        sel = np.concatenate([sel, np.argwhere(cost[0] == 255)[:1]])      # impassable target: empty frontier
    reqs = np.concatenate([capi.tile_req((0, 0), (int(r), int(c))) for r, c in sel])
    return p, cost, reqs


def portal_specs(portals, liid, cw, limit=None):
    """TARGET_PORTAL request specs for every portal: (chunk, idx, ep, next_chunk, next_ep, port_iid, next_iid)"""
    out = []
    key = portals[:, 0] * cw + portals[:, 1]
    for row in portals:
        cr, cc, idx, r0, c0, r1, c1, conn_chunk, conn_idx = [int(v) for v in row[:9]]
        nrow = portals[(key == conn_chunk) & (portals[:, 2] == conn_idx)][0]
        ncr, ncc = int(nrow[0]), int(nrow[1])
        nr0, nc0, nr1, nc1 = [int(v) for v in nrow[3:7]]
        piids = np.unique(liid[cr * cw + cc][r0:r1 + 1, c0:c1 + 1]); piids = piids[piids != 0xFFFF]
        niids = np.unique(liid[ncr * cw + ncc][nr0:nr1 + 1, nc0:nc1 + 1]); niids = niids[niids != 0xFFFF]
        for pi in list(piids[:2]) + [0xFFFF]:
            for ni in niids[:2]:
                out.append(((cr, cc), idx, (r0, c0, r1, c1), (ncr, ncc), (nr0, nc0, nr1, nc1), int(pi), int(ni)))
    return out[:limit] if limit else out


def portal_reqs(specs, init=1):
    return np.concatenate([capi.portal_req(s[0], s[2], s[3], s[4], s[5], s[6], init=init) for s in specs])


def los_case(cost, cw, ch, seed, ntargets=4):
    """LOS request batch: destination chunk + chained neighbours (+ one more hop). -> reqs"""
    rng = np.random.default_rng(seed)
    reqs = []
    for t in range(ntargets):
        chunk = (int(rng.integers(0, ch)), int(rng.integers(0, cw)))
        tiles = np.argwhere(cost[chunk[0] * cw + chunk[1]] != 255)
        tr, tc = [int(v) for v in tiles[rng.integers(0, len(tiles))]]
        td = (chunk[0], chunk[1], tr, tc)
        i0 = len(reqs)
        reqs.append(capi.los_req(chunk, td))
        for nb in ((chunk[0] + 1, chunk[1]), (chunk[0] - 1, chunk[1]), (chunk[0], chunk[1] + 1), (chunk[0], chunk[1] - 1)):
            if not (0 <= nb[0] < ch and 0 <= nb[1] < cw):
                continue
            i1 = len(reqs)
            reqs.append(capi.los_req(nb, td, prev_index=i0, prev_chunk=chunk))
            for nb2 in ((nb[0] + 1, nb[1]), (nb[0], nb[1] + 1), (nb[0] - 1, nb[1]), (nb[0], nb[1] - 1)):
                if not (0 <= nb2[0] < ch and 0 <= nb2[1] < cw) or nb2 == chunk:
                    continue
                reqs.append(capi.los_req(nb2, td, prev_index=i1, prev_chunk=nb))
    return np.concatenate(reqs)


def ref_los_batch(ref, reqs):
    """run a LOS request batch through the compiled reference"""
    out = np.zeros((len(reqs), 64, 64), np.uint8)
    for i, q in enumerate(reqs):
        td = (int(q["tgt_chunk_r"]), int(q["tgt_chunk_c"]), int(q["tgt_tile_r"]), int(q["tgt_tile_c"]))
        chunk = (int(q["chunk_r"]), int(q["chunk_c"]))
        if q["prev_index"] < 0:
            out[i] = ref.los(chunk, td)
        else:
            out[i] = ref.los(chunk, td, prev=out[int(q["prev_index"])], prev_chunk=(int(q["prev_chunk_r"]), int(q["prev_chunk_c"])))
    return out


def agent_case(cw, n, nflocks, seed, dens, spacing):
    """-> pathable, cost, agents dict (with ~10% ARRIVED and ~10% slow movers)"""
    p = noise_map(cw, cw, seed, dens)
    cost = synth.cost_from_pathable(p, cw, cw)
    a = synth.make_agents(cost, cw, cw, n, nflocks, seed, radius=1.0, spacing=spacing)
    rng = np.random.default_rng(seed)
    st = a["state"].copy(); st[rng.random(n) < 0.1] = 2
    a["state"] = st
    slow = rng.random(n) < 0.1
    a["vel"][slow] *= np.float32(0.05)
    a["prev_pos"] = (a["pos"] - a["vel"]).astype(np.float32)
    return p, cost, a


def relerr(g, e):
    d = np.abs(g - e).max(axis=1)
    return d / np.maximum(np.abs(e).max(axis=1), 1e-3)


def route_pairs(cost, cw, ch, seed, n):
    """n seeded (src_xz, dst_xz) pairs on passable tile centres"""
    rng = np.random.default_rng(seed)
    img = synth.blocked_to_image(cost, cw, ch)
    pas = np.argwhere(img != 255)
    out = []
    for _ in range(n):
        (sr, sc), (dr, dc) = pas[rng.integers(0, len(pas), 2)]
        out.append(((float(-(sc + 0.5) * 4), float((sr + 0.5) * 4)), (float(-(dc + 0.5) * 4), float((dr + 0.5) * 4))))
    return out


def execute_route(flow_exec, los_exec, fr, fc, lr, lc):
    """run a route's requests in order (flow merges in place per chunk; LOS as one chained batch)
    -> {chunk: flow}, {chunk: los}"""
    fields = {}
    for k in range(len(fr)):
        base = fields.get(int(fc[k]))
        out = flow_exec(fr[k:k + 1], None if fr[k]["init"] else base[None])
        fields[int(fc[k])] = out[0]
    los = los_exec(lr) if len(lr) else np.zeros((0, 64, 64), np.uint8)
    return fields, {int(lc[k]): los[k] for k in range(len(lr))}


def tile_attr_case(cw, ch, seed, terrain=False):
    """int32[H32][W32][4] = {pathable, type, base_height, ramp_height} (struct tile, tile.h:101).
    terrain=False: every tile type / height combination at random (stress for n_set_cost_for_tile and
    the cliff rule); terrain=True: plateaus of different heights joined by ramps and corner tiles,
    mostly pathable, so that portals / islands downstream are non-trivial."""
    rng = np.random.default_rng(seed)
    H, W = ch * 32, cw * 32
    t = np.zeros((H, W, 4), np.int32)
    if not terrain:
        t[..., 0] = rng.random((H, W)) < 0.9
        t[..., 1] = np.where(rng.random((H, W)) < 0.6, 0, rng.integers(0, 13, (H, W)))
        t[..., 2] = rng.integers(-3, 4, (H, W))
        t[..., 3] = rng.integers(0, 4, (H, W))
        return t
    t[..., 0] = 1
    for _ in range(3 * cw * ch):                       # plateaus (cliffs all around) and lakes
        h, w = rng.integers(4, 14, 2)
        r, c = rng.integers(0, H - h), rng.integers(0, W - w)
        t[r:r + h, c:c + w, 2] = rng.integers(-2, 3)
    for _ in range(6 * cw * ch):                       # ramps / corner tiles sprinkled on the borders
        r, c = rng.integers(0, H), rng.integers(0, W)
        t[r, c, 1] = rng.integers(1, 13)
        t[r, c, 3] = rng.integers(1, 3)
    t[..., 0] &= (rng.random((H, W)) > 0.03).astype(np.int32)
    return t


def dir_quat(v):
    """dir_quat_from_velocity (movement.c:1411), float64 maths is fine for INPUT generation"""
    ang = np.arctan2(v[..., 1], v[..., 0]) - np.pi / 2
    q = np.zeros(v.shape[:-1] + (4,), np.float32)
    q[..., 1] = np.sin(ang / 2); q[..., 3] = np.cos(ang / 2)
    return q


def update_case(seed, hz):
    """Agents + movestate for the state-update pass (entity_compute_update, movement.c:2303):
    flock 0 aims at its own centre (arrivals, adjacent-arrived cascade, closest-pathable clause),
    flock 1 aims into an obstacle, random facings trip the heading gate, some histories are empty,
    a few agents are combat-held / garrisoned / have no desired velocity."""
    cw = 3
    p, cost, a = agent_case(cw, 1500, 3, seed, 0.03, 2.6)
    rng = np.random.default_rng(seed + 1000)
    n = len(a["radius"])
    img = synth.blocked_to_image(cost, cw, cw)
    # flock 0 -> centre of its own disc; flock 1 -> the impassable tile nearest to its disc centre
    c0 = a["pos"][a["flock_of"] == 0].mean(axis=0)
    a["flock_target"][0] = c0
    c1 = a["pos"][a["flock_of"] == 1].mean(axis=0)
    imp = np.argwhere(img == 255)
    if len(imp):
        xz = np.stack([-(imp[:, 1] + 0.5) * 4.0, (imp[:, 0] + 0.5) * 4.0], axis=1)
        a["flock_target"][1] = xz[np.argmin(((xz - c1) ** 2).sum(axis=1))]
    tgt = a["flock_target"][a["flock_of"]]
    d = tgt - a["pos"]
    dist = np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-3)
    a["vel"] = (d / dist * (0.5 * a["max_speed"][:, None] / hz)).astype(np.float32)
    a["vel"][rng.random(n) < 0.1] *= np.float32(0.05)
    a["vel"][rng.random(n) < 0.05] = 0.0
    a["prev_pos"] = (a["pos"] - a["vel"]).astype(np.float32)
    # ARRIVED members start the adjacent-arrived cascade (movement.c:2457): many in flock 0, none in 1, few in 2
    st = np.zeros(n, np.int32)
    u = rng.random(n)
    st[(a["flock_of"] == 0) & (u < 0.05)] = 2
    st[(a["flock_of"] == 2) & (u < 0.004)] = 2
    a["state"] = st
    fl = a["flags"].copy()
    fl[rng.random(n) < 0.03] |= capi.FLAG_COMBAT_HELD
    fl[rng.random(n) < 0.01] |= capi.FLAG_GARRISONED
    a["flags"] = fl
    ms = np.zeros(n, capi.MOVESTATE)
    ms["next_pos"][:, 0] = a["pos"][:, 0]; ms["next_pos"][:, 2] = a["pos"][:, 1]
    ms["step"] = 1.0 / (20 // hz)
    ms["left"] = 0 if hz == 20 else (rng.random(n) < 0.3).astype(np.int32)
    face = np.where(np.linalg.norm(a["vel"], axis=1, keepdims=True) > 1e-3, a["vel"], d / dist)
    ang = rng.uniform(-2.4, 2.4, n) * (rng.random(n) < 0.5)
    rot = np.stack([face[:, 0] * np.cos(ang) - face[:, 1] * np.sin(ang), face[:, 0] * np.sin(ang) + face[:, 1] * np.cos(ang)], axis=1)
    ms["next_rot"] = dir_quat(rot)
    ms["combat_facing"] = dir_quat(rng.normal(size=(n, 2)))
    hist = np.repeat(a["vel"][:, None, :], 14, axis=1) + rng.normal(scale=0.02, size=(n, 14, 2))
    hist[rng.random(n) < 0.3] = 0.0
    ms["vel_hist"] = hist.astype(np.float32)
    ms["vel_hist_idx"] = rng.integers(0, 14, n)
    return p, cost, a, ms


def repair_case(ref, cw, ch, seed, per_chunk=4):
    """Requests for the repair chain of N_DesiredPointSeekVelocity (nav.c:3508-3554) on a map with blockers.
    ref: a pfref.RefMap that already holds the blockers (after update()). Returns (targets FIELD_REQ[n],
    kinds, args, base fields[n,64,64], expected[n,64,64])."""
    rng = np.random.default_rng(seed)
    cost, blk, liid = ref.cost_base(), ref.blockers(), ref.local_islands()
    ports = ref.portals()
    T, K, A, B, E = [], [], [], [], []
    for chunk in range(cw * ch):
        cr, cc = chunk // cw, chunk % cw
        npass = np.argwhere((cost[chunk] != 255) & (blk[chunk] == 0))
        nonp = np.argwhere((cost[chunk] == 255) | (blk[chunk] > 0))
        if len(npass) == 0:
            continue
        bases = []
        t = npass[rng.integers(len(npass))]
        bases.append((capi.tile_req((cr, cc), (int(t[0]), int(t[1]))), ref.flow_tile((cr, cc), (int(t[0]), int(t[1]))),
                      dict(tile=(int(t[0]), int(t[1])))))
        # a blocked target tile: the frontier is empty until `ignoreblock` (field.c:2367)
        bt = np.argwhere((cost[chunk] != 255) & (blk[chunk] > 0))
        if len(bt):
            t = bt[rng.integers(len(bt))]
            bases.append((capi.tile_req((cr, cc), (int(t[0]), int(t[1]))), ref.flow_tile((cr, cc), (int(t[0]), int(t[1]))),
                          dict(tile=(int(t[0]), int(t[1])))))
        specs = portal_specs(ports[(ports[:, 0] == cr) & (ports[:, 1] == cc)][:2] if False else ports, liid, cw)
        specs = [s for s in specs if s[0] == (cr, cc)][:3]
        for s in specs:
            bases.append((portal_reqs([s]), ref.flow_portal(s[0], s[1], s[5], s[6]), dict(portal=(s[1], s[5], s[6]))))
        iids = [int(i) for i in np.unique(liid[chunk]) if i != 0xFFFF]
        for req, base, kw in bases:
            for s in (nonp[rng.integers(0, len(nonp), per_chunk)] if len(nonp) else []):
                T.append(req); K.append(0); A.append((int(s[0]) << 8) | int(s[1])); B.append(base)
                E.append(ref.flow_nearest_pathable((cr, cc), (int(s[0]), int(s[1])), base))
            for iid in iids[:per_chunk + 2]:
                T.append(req); K.append(1); A.append(iid); B.append(base)
                E.append(ref.flow_island_to_nearest((cr, cc), iid, base, **kw))
    return (np.concatenate(T), np.array(K, np.int32), np.array(A, np.int32), np.stack(B), np.stack(E))


def region_case(seed, cw=3, ch=3, n=40, dim=96):
    """Region-field scenario ("cell arrival" fields, field.c:2445-2711) on a map with faction blockers.
    -> (pathable, blockers[(x, z, radius, faction)], wars, reqs) where reqs is a list of dicts
    {center, target, enemies, overlay, start} in absolute tile coordinates (start = None: no fixup pass).
    Covers: interior regions, regions hanging over every map edge (incl. the rows < columns clamp whose visited
    array aliases, field.c:1431), a target one past the far edge (base shift, field.c:2477), overlay tiles,
    enemy masks and blocked / impassable fixup starts."""
    rng = np.random.default_rng(seed)
    p = noise_map(cw, ch, seed, 0.10)
    H, W = ch * 64, cw * 64
    blockers = [(float(-rng.uniform(10, cw * 256 - 10)), float(rng.uniform(10, ch * 256 - 10)), float(rng.uniform(2, 12)),
                 int(rng.integers(0, 4))) for _ in range(80)]
    wars = [(0, 1), (0, 2), (3, 1)]
    half = dim // 2
    reqs = []
    for i in range(n):
        kind = i % 8
        if kind in (0, 1, 2):      # interior
            cr_, cc_ = int(rng.integers(half, H - half)), int(rng.integers(half, W - half))
        elif kind == 3:            # top edge (rows clamped, columns not: aliasing flood)
            cr_, cc_ = int(rng.integers(0, half - 4)), int(rng.integers(half, W - half))
        elif kind == 4:            # left / bottom
            cr_, cc_ = int(rng.integers(H - half + 2, H)), int(rng.integers(0, half))
        elif kind == 5:            # right edge
            cr_, cc_ = int(rng.integers(half, H - half)), int(rng.integers(W - half + 1, W))
        elif kind == 6:            # corner
            cr_, cc_ = int(rng.integers(0, 20)), int(rng.integers(W - 20, W))
        else:
            cr_, cc_ = int(rng.integers(0, H)), int(rng.integers(0, W))
        lo_r, hi_r = max(cr_ - half, 0), min(cr_ + half - 1, H - 1)
        lo_c, hi_c = max(cr_ * 0 + cc_ - half, 0), min(cc_ + half - 1, W - 1)
        tr, tc = int(rng.integers(lo_r, hi_r + 1)), int(rng.integers(lo_c, hi_c + 1))
        if i % 11 == 5 and cr_ + half < H and cc_ + half < W:
            tr, tc = cr_ + half, cc_ + half          # one past the far edge: the base shifts
        ov = None
        if i % 3 == 1:
            k = int(rng.integers(1, 40))
            ov = np.stack([rng.integers(cr_ - half - 3, cr_ + half + 3, k), rng.integers(cc_ - half - 3, cc_ + half + 3, k)], 1)
            ov = ov[(ov[:, 0] >= 0) & (ov[:, 0] < H) & (ov[:, 1] >= 0) & (ov[:, 1] < W)].astype(np.int32)
        enemies = [0, 0, 0b0010, 0b0110, 0b1111][i % 5]
        reqs.append(dict(center=(cr_, cc_), target=(tr, tc), enemies=enemies, overlay=ov, start=None, want_fixup=(i % 2 == 0)))
    return p, blockers, wars, reqs


def region_pick_starts(reqs, cost, blk, cw, ch, seed, dim=96):
    """choose the fixup start of every request that wants one: a non-passable tile inside the clamped region
    (clamped_region, field.c:1892), nearest to the centre in a seeded scan order"""
    rng = np.random.default_rng(seed)
    H, W = ch * 64, cw * 64
    half = dim // 2
    img_c = cost.reshape(ch, cw, 64, 64).transpose(0, 2, 1, 3).reshape(H, W)
    img_b = blk.reshape(ch, cw, 64, 64).transpose(0, 2, 1, 3).reshape(H, W)
    nonp = (img_c == 255) | (img_b > 0)
    for q in reqs:
        if not q["want_fixup"]:
            continue
        cr_, cc_ = q["center"]
        b_r = cr_ - half if cr_ - half >= 0 else 0
        b_c = cc_ - half if cc_ - half >= 0 else 0
        e_r = cr_ + half if cr_ + half < H else H - 1
        e_c = cc_ + half if cc_ + half < W else W - 1
        cand = np.argwhere(nonp[b_r:e_r, b_c:e_c])
        if len(cand) == 0:
            continue
        s = cand[rng.integers(len(cand))]
        q["start"] = (int(s[0]) + b_r, int(s[1]) + b_c)
    return reqs


def region_reqs_from_golden(rec, ov):
    """inverse of make_golden.pack_region_reqs"""
    out, o = [], 0
    for row in rec:
        n = int(row[7])
        out.append(dict(center=(int(row[0]), int(row[1])), target=(int(row[2]), int(row[3])), enemies=int(row[4]),
                        start=None if row[5] < 0 else (int(row[5]), int(row[6])), overlay=ov[o:o + n] if n else None))
        o += n
    return out


def target_case(seed, cw=3, ch=3, n=120):
    """entities for the TARGET_ENTITY / TARGET_ENEMIES fields: positions, selection radii, factions, flags
    (all MOVABLE, ~85 % COMBATABLE) -> dict"""
    rng = np.random.default_rng(seed)
    pos = np.stack([-rng.uniform(4, cw * 256 - 4, n), rng.uniform(4, ch * 256 - 4, n)], 1).astype(np.float32)
    radius = rng.uniform(0.5, 9.0, n).astype(np.float32)
    factions = rng.integers(0, 4, n).astype(np.int32)
    flags = np.full(n, (1 << 3) | (1 << 4), np.uint32)
    flags[rng.random(n) < 0.15] &= ~np.uint32(1 << 4)
    return dict(pos=pos, radius=radius, factions=factions, flags=flags)


def enemies_of(faction, wars, factions, flags):
    """field_enemy_ent (field.c:963) with fog disabled: other faction, COMBATABLE, at war"""
    foes = [b if a == faction else a for a, b in wars if faction in (a, b)]
    return np.isin(factions, foes) & ((flags & (1 << 4)) != 0)


def stress_layout(army=256, spacing=12.0):
    """the unit layout of the reference's own stress test (scripts/test_stress.py:73-121): two armies of `army` units in
    4 rows on the 4 x 4-chunk plain map centred at the origin, red (selection radius 3.25) attacking (-100, 0), blue
    (3.00) attacking (+100, 0). -> pos[2 * army, 2], radius, flock_of, targets[2, 2]"""
    import math
    nrows = 4; ncols = math.ceil(army / nrows)
    red, blue = [], []
    for r in range(int(-nrows // 2), int(nrows // 2 + nrows % 2)):
        for c in range(int(-ncols // 2), int(ncols // 2 + ncols % 2)):
            red.append((-(r * spacing) + 35.0, c * spacing))
            blue.append(((r * spacing) - 35.0, c * spacing))
    pos = np.array(red + blue, np.float32)
    radius = np.concatenate([np.full(len(red), 3.25, np.float32), np.full(len(blue), 3.0, np.float32)])
    flock_of = np.concatenate([np.zeros(len(red), np.int32), np.ones(len(blue), np.int32)])
    return pos, radius, flock_of, np.array([[-100.0, 0.0], [100.0, 0.0]], np.float32)
