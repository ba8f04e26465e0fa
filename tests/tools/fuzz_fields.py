"""offline fuzz: flow (tile + portal targets, factions) and chained LOS fields, port vs the compiled reference, on maps
with faction blockers"""
import os, sys, time, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import cases, pfref, pforacle
capi = cases.capi
t0 = time.time(); tot = bad = 0
for seed in range(2000, 2010):
    cw, ch = [(2, 2), (3, 2), (2, 3), (3, 3)][seed % 4]
    rng = np.random.default_rng(seed)
    p = cases.synth.make_map(cw, ch, seed, frac_blocked=0.12, rivers=True); p[rng.random(p.shape) < [0.03, 0.12, 0.25][seed % 3]] = 0
    ref = pfref.RefMap(cw, ch, p)
    wars = [(0, 1), (0, 2), (3, 1), (2, 3)][: 2 + seed % 3]
    for a, b in wars: ref.set_war(a, b)
    for _ in range(50):
        ref.blockers_incref(float(-rng.uniform(10, cw * 256 - 10)), float(rng.uniform(10, ch * 256 - 10)), float(rng.uniform(2, 10)), int(rng.integers(0, 4)), 0)
    ref.update()
    cost, blk, liid, fac = ref.cost_base(), ref.blockers(), ref.local_islands(), ref.factions()
    enemies = np.zeros(16, np.uint16)
    for a, b in wars: enemies[a] |= 1 << b; enemies[b] |= 1 << a
    om = pforacle.OracleMap(cw, ch, cost, blk, liid, factions=fac, enemies=enemies)
    ports = ref.portals()
    specs = cases.portal_specs(ports, liid, cw, limit=30)
    b = n = 0
    for f in (0xF, 0, 1, 3):
        for chunk in range(cw * ch):
            cr, cc = chunk // cw, chunk % cw
            npass = np.argwhere(cost[chunk] != 255)
            if len(npass) == 0: continue
            for t in npass[rng.integers(0, len(npass), 3)]:
                q = capi.tile_req((cr, cc), (int(t[0]), int(t[1]))); q["faction_id"] = f
                e = ref.flow_tile((cr, cc), (int(t[0]), int(t[1])), faction=f)
                n += 1; b += int((om.flow_fields_update(q)[0] != e).any())
                td = (cr, cc, int(t[0]), int(t[1]))
                ql = capi.los_req((cr, cc), td); ql["faction_id"] = f
                l0 = ref.los((cr, cc), td, faction=f)
                g0 = om.los_fields_create(ql)[0]
                n += 1; b += int((g0 != l0).any())
                for nb in ((cr, cc + 1), (cr + 1, cc), (cr, cc - 1), (cr - 1, cc)):
                    if not (0 <= nb[0] < ch and 0 <= nb[1] < cw): continue
                    qn = np.concatenate([ql, capi.los_req(nb, td, prev_index=0, prev_chunk=(cr, cc))]); qn["faction_id"] = f
                    e1 = ref.los(nb, td, prev=l0, prev_chunk=(cr, cc), faction=f)
                    n += 1; b += int((om.los_fields_create(qn)[1] != e1).any())
        for s_ in specs[:12]:
            q = cases.portal_reqs([s_]); q["faction_id"] = f
            n += 1; b += int((om.flow_fields_update(q)[0] != ref.flow_portal(s_[0], s_[1], s_[5], s_[6], faction=f)).any())
    tot += n; bad += b
    print("seed", seed, (cw, ch), "fields", n, "bad", b, "%.0fs" % (time.time() - t0), flush=True)
    ref.close()
print("TOTAL", tot, "bad", bad)
