"""offline fuzz: entity_compute_update, port vs the compiled reference, at 20 / 10 / 5 / 1 Hz on new populations"""
import os, sys, time, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import cases, pfref, pforacle
capi = cases.capi
t0 = time.time()
for seed, hz in ((81, 5), (82, 1), (83, 20), (84, 10), (85, 5), (86, 1)):
    cw = 3
    p, cost, a, ms = cases.update_case(seed, hz)
    nflocks = len(a["flock_target"])
    ref = pfref.RefMap(cw, cw, p)
    dest_ids = []
    for f in range(nflocks):
        src = a["pos"][np.argmax(a["flock_of"] == f)]; tgt = a["flock_target"][f]
        ok, did = ref.request_path((float(src[0]), float(src[1])), (float(tgt[0]), float(tgt[1])))
        dest_ids.append(did if ok else ref.dest_id((float(tgt[0]), float(tgt[1]))))
    ref.agents_set(a["pos"], a["prev_pos"], a["vel"], a["radius"], a["max_speed"], a["state"], a["flags"],
                   a["flock_of"], a["flock_target"], np.array(dest_ids, np.uint32), hz=hz)
    ref.movestate_set(ms["next_pos"][:, [0, 2]], ms["next_rot"], ms["step"], ms["left"], ms["vel_hist"],
                      ms["vel_hist_idx"], np.zeros(len(ms), np.int32), np.zeros(len(ms), np.int32), ms["combat_facing"])
    work = np.nonzero((a["state"] != 2) & (a["state"] != 4))[0].astype(np.uint32)
    vdes = np.zeros((len(work), 2), np.float32); los = np.zeros(len(work), np.uint8)
    for _pass in range(2):
        for f in range(nflocks):
            sel = np.nonzero(a["flock_of"][work] == f)[0]
            if len(sel) == 0: continue
            v, l = ref.desired_velocity(dest_ids[f], a["pos"][work[sel]], a["prev_pos"][work[sel]], a["flock_target"][f])
            vdes[sel] = v; los[sel] = l
    rng = np.random.default_rng(seed)
    vdes[rng.random(len(work)) < 0.03] = 0.0
    ref.work_set(work, vdes, los, a["speed"][work])
    vel, _ = ref.velocity_work(1)
    oi, of = ref.compute_updates(vel)
    n = len(a["radius"])
    a2 = dict(a); a2["vdes"] = np.zeros((n, 2), np.float32); a2["vdes"][work] = vdes
    a2["has_los"] = np.zeros(n, np.uint32); a2["has_los"][work] = los
    rec, fl = capi.pack_agents(a2)
    nav = capi.Nav(hostonly=True)
    nav.map_create(cw, cw, 1); nav.map_upload_layer(0, cost); nav.map_build_nav(0); nav.route_build(0)
    arrival = [nav.route_arrival_consts(t) for t in a["flock_target"]]
    nav.close()
    w = pforacle.OracleWorld(pforacle.OracleMap(cw, cw, cost), rec, fl, hz)
    pp = w.entity_updates(ms, arrival, work, vel, vdes, capi.PATCH)
    same = (pp["flags"] == oi[:, 0].astype(np.uint32)) & (pp["next_state"] == oi[:, 1]) & (pp["next_block"] == oi[:, 2])
    got = np.concatenate([pp["next_velocity"], pp["next_pos"], pp["next_rot"], pp["next_ppos"], pp["next_npos"], pp["next_step"][:, None],
                          pp["next_left"][:, None], pp["next_nrot"], pp["next_prot"]], axis=1)
    print("seed", seed, "hz", hz, "work", len(work), "discrete mismatches", int((~same).sum()), "float mismatches", int((got[same] != of[same][:, :25]).any(axis=1).sum()),
          "states", {int(k): int(v) for k, v in zip(*np.unique(oi[:, 1], return_counts=True))}, "%.0fs" % (time.time() - t0), flush=True)
    w.close(); ref.close()
