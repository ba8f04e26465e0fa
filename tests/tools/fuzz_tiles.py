"""offline fuzz: the tile -> cost pass (all 12 reference layers) and OBB blockers, port / host code vs the compiled reference"""
import os, sys, time, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import cases, pfref, pforacle
capi = cases.capi
t0 = time.time(); tot = bad = 0
for seed in range(3000, 3008):
    cw, ch = [(2, 2), (3, 2), (1, 3), (4, 1)][seed % 4]
    t = cases.tile_attr_case(cw, ch, seed, terrain=bool(seed % 2))
    ref = pfref.RefMap(cw, ch, tiles=t)
    b = 0
    for L in range(12):
        b += int((pforacle.cost_from_tiles(cw, ch, t, L) != ref.cost_base(L)).any())
    # OBB + circle blockers on the four ground layers through the host code
    nav = capi.Nav(hostonly=True)
    nav.map_create(cw, ch, 4)
    for L in range(4):
        nav.map_upload_layer(L, ref.cost_base(L)); nav.map_build_nav(L)
    rng = np.random.default_rng(seed)
    for _ in range(25):
        c = np.array([-rng.uniform(45, cw * 256 - 45), rng.uniform(45, ch * 256 - 45)])
        ang, hx, hz = rng.uniform(0, np.pi), rng.uniform(2, 40), rng.uniform(2, 40)
        ax, az = np.array([np.cos(ang), np.sin(ang)]), np.array([-np.sin(ang), np.cos(ang)])
        corners = np.array([c - ax * hx - az * hz, c + ax * hx - az * hz, c + ax * hx + az * hz, c - ax * hx + az * hz], np.float32)
        inside = (corners[:, 0] < -0.5).all() and (corners[:, 0] > -(cw * 256 - 0.5)).all() and (corners[:, 1] > 0.5).all() and (corners[:, 1] < ch * 256 - 0.5).all()
        if not inside: continue
        f = int(rng.integers(0, 5))
        ref.blockers_obb(corners, True, f, 0); nav.blockers_obb(corners, True, f, 0)
    ref.update(); nav.map_commit()
    bb = 0
    for L in range(4):
        bb += int((nav.blockers(L) != ref.blockers(L)).any()) + int((nav.local_islands(L) != ref.local_islands(L)).any()) + int((nav.faction_counts(L) != ref.factions(L)).any())
    tot += 12 + 12; bad += b + bb
    print("seed", seed, (cw, ch), "cost layers bad", b, "of 12 | blocker/island/faction arrays bad", bb, "of 12  %.0fs" % (time.time() - t0), flush=True)
    ref.close(); nav.close()
print("TOTAL", tot, "bad", bad)
