"""offline fuzz: n_request_path (host planner + port fields) vs the compiled reference on more maps, incl. tile-attribute
terrain maps (ramps / cliffs), different densities and map shapes"""
import os, sys, numpy as np, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import cases, pfref, pforacle
import test_oracle as T
capi = cases.capi
def run(cw, ch, ref, seed, npairs):
    cost = ref.cost_base()
    nav = capi.Nav(hostonly=True)
    nav.map_create(cw, ch, 1); nav.map_upload_layer(0, cost); nav.map_build_nav(0); nav.route_build(0)
    assert (nav.local_islands(0) == ref.local_islands()).all()
    assert (nav.route_islands(0) == ref.islands()).all()
    assert (nav.portals(0)[:, :9] == ref.portals()[:, :9]).all()
    om = pforacle.OracleMap(cw, ch, cost, None, ref.local_islands())
    pairs = cases.route_pairs(cost, cw, ch, seed, npairs)
    oks, dids, ffids, flows, loss, has = [], [], [], [], [], []
    for src, dst in pairs:
        ref.fc_clear()
        ok, did = ref.request_path(src, dst)
        oks.append(ok); dids.append(did)
        fid = np.zeros(cw * ch, np.uint64); hs = np.zeros(cw * ch, np.uint8)
        fl = np.zeros((cw * ch, 64, 64), np.uint8); ls = np.zeros((cw * ch, 64, 64), np.uint8)
        for c in range(cw * ch):
            f, i = ref.fc_flow(did, (c // cw, c % cw)) if ok else (None, None)
            l = ref.fc_los(did, (c // cw, c % cw)) if ok else None
            if f is not None: fl[c] = f; fid[c] = i; hs[c] |= 1
            if l is not None: ls[c] = l; hs[c] |= 2
        ffids.append(fid); flows.append(fl); loss.append(ls); has.append(hs)
    T._check_route_against(nav, om, cw, ch, pairs, oks, dids, ffids, flows, loss, has)
    nav.close()
    return sum(oks), len(oks)
t0 = time.time()
for seed, (cw, ch), dens in ((101, (4, 4), 0.05), (102, (5, 3), 0.2), (103, (2, 6), 0.35), (104, (6, 6), 0.12), (105, (3, 3), 0.45)):
    p = cases.noise_map(cw, ch, seed, dens) if cw == ch else None
    if p is None:
        rng = np.random.default_rng(seed)
        p = cases.synth.make_map(cw, ch, seed, frac_blocked=0.12, rivers=True)
        p[rng.random(p.shape) < dens] = 0
    ref = pfref.RefMap(cw, ch, p)
    print("noise", seed, (cw, ch), dens, run(cw, ch, ref, seed, 40), "%.0fs" % (time.time() - t0), flush=True)
    ref.close()
for seed, (cw, ch) in ((201, (3, 3)), (202, (4, 2)), (203, (4, 4))):
    t = cases.tile_attr_case(cw, ch, seed, terrain=True)
    ref = pfref.RefMap(cw, ch, tiles=t)
    print("terrain", seed, (cw, ch), run(cw, ch, ref, seed, 40), "%.0fs" % (time.time() - t0), flush=True)
    ref.close()
