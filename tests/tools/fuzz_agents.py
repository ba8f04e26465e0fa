"""offline fuzz: the port's velocity pass vs the compiled reference on new seeds / densities / radii"""
import os, sys, numpy as np, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import cases, pfref, pforacle
capi, synth = cases.capi, cases.synth
t0 = time.time()
for seed, cw, n, nflocks, dens, spacing, radius in ((301, 2, 900, 2, 0.04, 2.3, 1.0), (302, 1, 500, 1, 0.0, 2.05, 1.5), (303, 3, 1200, 4, 0.08, 3.0, 0.75),
                                                    (304, 2, 700, 3, 0.02, 2.2, 3.0), (305, 1, 350, 2, 0.1, 4.0, 1.0)):
    p = cases.noise_map(cw, cw, seed, dens)
    cost = synth.cost_from_pathable(p, cw, cw)
    a = synth.make_agents(cost, cw, cw, n, nflocks, seed, radius=radius, spacing=spacing)
    rng = np.random.default_rng(seed)
    st = a["state"].copy(); st[rng.random(n) < 0.1] = 2; a["state"] = st
    a["prev_pos"] = (a["pos"] - a["vel"]).astype(np.float32)
    ref = pfref.RefMap(cw, cw, p)
    dest_ids = []
    for f in range(nflocks):
        src = a["pos"][np.argmax(a["flock_of"] == f)]; tgt = a["flock_target"][f]
        ok, did = ref.request_path((float(src[0]), float(src[1])), (float(tgt[0]), float(tgt[1])))
        dest_ids.append(did if ok else ref.dest_id((float(tgt[0]), float(tgt[1]))))
    ref.agents_set(a["pos"], a["prev_pos"], a["vel"], a["radius"], a["max_speed"], a["state"], a["flags"],
                   a["flock_of"], a["flock_target"], np.array(dest_ids, np.uint32), hz=20)
    work = np.nonzero((a["state"] != 2) & (a["state"] != 4))[0].astype(np.uint32)
    vdes = np.zeros((len(work), 2), np.float32); los = np.zeros(len(work), np.uint8)
    for _pass in range(2):
        for f in range(nflocks):
            sel = np.nonzero(a["flock_of"][work] == f)[0]
            if len(sel) == 0: continue
            v, l = ref.desired_velocity(dest_ids[f], a["pos"][work[sel]], a["prev_pos"][work[sel]], a["flock_target"][f])
            vdes[sel] = v; los[sel] = l
    ref.work_set(work, vdes, los, a["speed"][work])
    vel, _ = ref.velocity_work(1)
    vpref = ref.vpref()
    a2 = dict(a); a2["vdes"] = np.zeros((n, 2), np.float32); a2["vdes"][work] = vdes
    a2["has_los"] = np.zeros(n, np.uint32); a2["has_los"][work] = los
    rec, fl = capi.pack_agents(a2)
    om = pforacle.OracleMap(cw, cw, cost)
    w = pforacle.OracleWorld(om, rec, fl, 20)
    pv, pp = w.velocity_work(work)
    ev, ep = cases.relerr(pv, vel), cases.relerr(pp, vpref)
    print(seed, "agents", n, "work", len(work), "vel relerr max %.2e (>1e-4: %d)  vpref relerr max %.2e  exact %.1f%%  %.0fs" % (
        ev.max(), int((ev > 1e-4).sum()), ep.max(), 100.0 * (pv == vel).all(axis=1).mean(), time.time() - t0), flush=True)
    w.close(); ref.close()
