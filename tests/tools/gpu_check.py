#!/usr/bin/env python3
"""Developer check (GPU box): CUDA path vs the compiled reference (oracle/_ref) on seeded cases.
Prints mismatch statistics; exits non-zero on any integer mismatch."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
pf = importlib.import_module("permafrost-engine_b200")
import pfref
capi, synth = pf.capi, pf.synth

fails = 0
def check(name, ok, extra=""):
    global fails
    print(("PASS " if ok else "FAIL ") + name + " " + extra, flush=True)
    if not ok: fails += 1

def noise_map(cw, ch, seed, density):
    rng = np.random.default_rng(seed)
    p = synth.make_map(cw, ch, seed, frac_blocked=0.12, rivers=(cw*ch > 1))
    if density > 0:
        p[rng.random(p.shape) < density] = 0
    return p

def test_flow_tile(nav, tma):
    nav.set_tma(tma)
    for seed, dens in ((1, 0.0), (2, 0.25), (3, 0.05), (4, 0.4)):
        p = noise_map(1, 1, seed, dens)
        ref = pfref.RefMap(1, 1, p)
        cost = synth.cost_from_pathable(p, 1, 1)
        assert (cost == ref.cost_base()).all()
        nav.map_create(1, 1, 1)
        nav.map_upload_layer(0, cost)
        rng = np.random.default_rng(seed)
        tiles = np.argwhere(cost[0] != 255)
        sel = tiles[rng.integers(0, len(tiles), 24)]
        sel = np.concatenate([sel, np.argwhere(cost[0] == 255)[:2]]) if (cost[0] == 255).any() else sel
        reqs = np.concatenate([capi.tile_req((0, 0), (int(r), int(c))) for r, c in sel])
        got = nav.flow_fields_update(reqs)
        bad = 0
        for k, (r, c) in enumerate(sel):
            exp = ref.flow_tile((0, 0), (int(r), int(c)))
            bad += int((exp != got[k]).sum())
        check(f"flow TILE tma={tma} seed={seed} dens={dens}", bad == 0, f"mismatched tiles={bad}")
        ref.close()

def test_flow_general(nav):
    import torch
    p = noise_map(1, 1, 7, 0.2)
    ref = pfref.RefMap(1, 1, p)
    cost = synth.cost_from_pathable(p, 1, 1)
    nav.map_create(1, 1, 1); nav.map_upload_layer(0, cost)
    tiles = np.argwhere(cost[0] != 255)[::97][:16]
    reqs = np.concatenate([capi.tile_req((0, 0), (int(r), int(c))) for r, c in tiles])
    d_reqs = torch.from_numpy(reqs.view(np.uint8)).cuda()
    d_out = torch.zeros((len(reqs), 4096), dtype=torch.uint8, device="cuda")
    nav.flow_fields_update_dev(d_reqs.data_ptr(), len(reqs), d_out.data_ptr(), 0, general=True)
    torch.cuda.synchronize()
    got = d_out.cpu().numpy().reshape(-1, 64, 64)
    bad = sum(int((ref.flow_tile((0, 0), (int(r), int(c))) != got[k]).sum()) for k, (r, c) in enumerate(tiles))
    check("flow general-cost kernel (unit map)", bad == 0, f"mismatched tiles={bad}")
    ref.close()

def portal_reqs_from_ref(ref, layer=0):
    """every (portal -> connected portal) TARGET_PORTAL request with the ISLAND ids seen on the portal tiles"""
    ports = ref.portals(layer)
    liid = ref.local_islands(layer)
    cw = ref.cw
    out = []
    for row in ports:
        cr, cc, idx, r0, c0, r1, c1, conn_chunk, conn_idx, nn = [int(v) for v in row]
        nrow = ports[(ports[:, 0] * cw + ports[:, 1] == conn_chunk) & (ports[:, 2] == conn_idx)][0]
        ncr, ncc = int(nrow[0]), int(nrow[1])
        nr0, nc0, nr1, nc1 = [int(v) for v in nrow[3:7]]
        piids = np.unique(liid[cr * cw + cc][r0:r1 + 1, c0:c1 + 1]); piids = piids[piids != 0xFFFF]
        niids = np.unique(liid[ncr * cw + ncc][nr0:nr1 + 1, nc0:nc1 + 1]); niids = niids[niids != 0xFFFF]
        for pi in list(piids[:2]) + [0xFFFF]:
            for ni in niids[:2]:
                out.append(((cr, cc), idx, (r0, c0, r1, c1), (ncr, ncc), (nr0, nc0, nr1, nc1), int(pi), int(ni)))
    return out

def test_flow_portal(nav, tma):
    nav.set_tma(tma)
    for seed, dens in ((11, 0.0), (12, 0.15)):
        cw = ch = 3
        p = noise_map(cw, ch, seed, dens)
        ref = pfref.RefMap(cw, ch, p)
        cost = ref.cost_base(); liid = ref.local_islands()
        nav.map_create(cw, ch, 1); nav.map_upload_layer(0, cost, None, liid)
        specs = portal_reqs_from_ref(ref)[:160]
        reqs = np.concatenate([capi.portal_req(s[0], s[2], s[3], s[4], s[5], s[6]) for s in specs])
        got = nav.flow_fields_update(reqs)
        bad = 0
        for k, s in enumerate(specs):
            exp = ref.flow_portal(s[0], s[1], s[5], s[6])
            bad += int((exp != got[k]).sum())
        check(f"flow PORTAL tma={tma} seed={seed} n={len(specs)}", bad == 0, f"mismatched tiles={bad}")
        # in-place update semantics (nav.c:1998-2008): second target merged into an existing field
        s0, s1 = specs[0], specs[-1]
        base = ref.flow_tile(s0[0], (5, 5))
        exp = ref.flow_portal(s0[0], s0[1], s0[5], s0[6], inout=base)
        q = capi.portal_req(s0[0], s0[2], s0[3], s0[4], s0[5], s0[6], init=0)
        got2 = nav.flow_fields_update(q, inout=base[None])
        check(f"flow PORTAL in-place update tma={tma} seed={seed}", (got2[0] == exp).all())
        ref.close()

def test_los(nav):
    for seed, dens in ((21, 0.0), (22, 0.08), (23, 0.2)):
        cw = ch = 2
        p = noise_map(cw, ch, seed, dens)
        ref = pfref.RefMap(cw, ch, p)
        cost = ref.cost_base()
        nav.map_create(cw, ch, 1); nav.map_upload_layer(0, cost)
        rng = np.random.default_rng(seed)
        reqs, exps = [], []
        for t in range(12):
            chunk = (int(rng.integers(0, ch)), int(rng.integers(0, cw)))
            tiles = np.argwhere(cost[chunk[0] * cw + chunk[1]] != 255)
            tr, tc = [int(v) for v in tiles[rng.integers(0, len(tiles))]]
            td = (chunk[0], chunk[1], tr, tc)
            e0 = ref.los(chunk, td)
            i0 = len(reqs)
            reqs.append(capi.los_req(chunk, td)); exps.append(e0)
            # neighbours chained off the destination chunk, then one more hop
            for nb in ((chunk[0] + 1, chunk[1]), (chunk[0] - 1, chunk[1]), (chunk[0], chunk[1] + 1), (chunk[0], chunk[1] - 1)):
                if not (0 <= nb[0] < ch and 0 <= nb[1] < cw): continue
                e1 = ref.los(nb, td, prev=e0, prev_chunk=chunk)
                i1 = len(reqs)
                reqs.append(capi.los_req(nb, td, prev_index=i0, prev_chunk=chunk)); exps.append(e1)
                for nb2 in ((nb[0] + 1, nb[1]), (nb[0], nb[1] + 1), (nb[0] - 1, nb[1]), (nb[0], nb[1] - 1)):
                    if not (0 <= nb2[0] < ch and 0 <= nb2[1] < cw) or nb2 == chunk: continue
                    e2 = ref.los(nb2, td, prev=e1, prev_chunk=nb)
                    reqs.append(capi.los_req(nb2, td, prev_index=i1, prev_chunk=nb)); exps.append(e2)
        got = nav.los_fields_create(np.concatenate(reqs))
        bad = sum(int((exps[k] != got[k]).sum()) for k in range(len(exps)))
        nbadf = sum(int((exps[k] != got[k]).any()) for k in range(len(exps)))
        check(f"LOS seed={seed} dens={dens} n={len(exps)}", bad == 0, f"mismatched tiles={bad} in {nbadf} fields")
        ref.close()

def setup_agents(cw, ch, n, nflocks, seed, dens, radius=1.0, spacing=2.6):
    p = noise_map(cw, ch, seed, dens)
    ref = pfref.RefMap(cw, ch, p)
    cost = ref.cost_base()
    a = synth.make_agents(cost, cw, ch, n, nflocks, seed, radius=radius, spacing=spacing)
    return p, ref, cost, a

def test_agents(nav, cw, n, nflocks, seed, dens, spacing, from_pool):
    ch = cw
    p, ref, cost, a = setup_agents(cw, ch, n, nflocks, seed, dens, spacing=spacing)
    nav.map_create(cw, ch, 1); nav.map_upload_layer(0, cost, None, ref.local_islands())
    rng = np.random.default_rng(seed)
    # a few agents become static obstacles (ARRIVED) / slow movers
    st = a["state"].copy(); st[rng.random(n) < 0.1] = 2
    a["state"] = st
    slow = rng.random(n) < 0.1
    a["vel"][slow] *= 0.05
    dest_ids = []
    for f in range(nflocks):
        tt = a["flock_target_tile"][f]
        # a representative source: the first agent of the flock
        src = a["pos"][np.argmax(a["flock_of"] == f)]
        ok, did = ref.request_path((float(src[0]), float(src[1])), (float(a["flock_target"][f][0]), float(a["flock_target"][f][1])))
        dest_ids.append(did if ok else ref.dest_id((float(a["flock_target"][f][0]), float(a["flock_target"][f][1]))))
    ref.agents_set(a["pos"], a["prev_pos"], a["vel"], a["radius"], a["max_speed"], a["state"], a["flags"],
                   a["flock_of"], a["flock_target"], np.array(dest_ids, np.uint32), hz=20)
    work = np.nonzero((a["state"] != 2) & (a["state"] != 4))[0].astype(np.uint32)
    # reference vdes / LOS per work item (per flock)
    vdes = np.zeros((len(work), 2), np.float32); los = np.zeros(len(work), np.uint8)
    # pass 0 warms the reference's field cache (N_DesiredPointSeekVelocity requests paths on a miss and
    # may merge further targets into a cached field, nav.c:1998-2008); pass 1 reads the settled state
    for _pass in range(2):
        for f in range(nflocks):
            sel = np.nonzero(a["flock_of"][work] == f)[0]
            if len(sel) == 0: continue
            v, l = ref.desired_velocity(dest_ids[f], a["pos"][work[sel]], a["prev_pos"][work[sel]], a["flock_target"][f])
            vdes[sel] = v; los[sel] = l
    ref.work_set(work, vdes, los, a["speed"][work])
    t0 = time.time(); exp_vel, secs = ref.velocity_work(1); t_ref = time.time() - t0
    exp_vpref = ref.vpref()

    # ---- spatial index order ----
    nav_agents = dict(a); nav_agents["vdes"] = np.zeros((n, 2), np.float32); nav_agents["has_los"] = np.zeros(n, np.uint32)
    nav_agents["vdes"][work] = vdes; nav_agents["has_los"][work] = los
    rec, fl = capi.pack_agents(nav_agents)
    nav.agents_upload(rec, fl, 20)
    badq = 0
    for i in rng.integers(0, n, 20):
        for r in (10.0, 30.0):
            e = ref.ents_in_circle(float(a["pos"][i, 0]), float(a["pos"][i, 1]), r, 512)
            g = nav.ents_in_circle(float(a["pos"][i, 0]), float(a["pos"][i, 1]), r, 512)
            if len(e) != len(g) or (e != g).any(): badq += 1
    check(f"agents[{cw}x{cw},n={n}] ents_in_circle order", badq == 0, f"bad queries={badq}")

    if from_pool:
        # fill the device pool with the reference's own cached fields for every (dest, chunk)
        nav.pool_create(nflocks, nflocks * cw * ch)
        for f in range(nflocks):
            for cr in range(ch):
                for cc in range(cw):
                    ff, _ = ref.fc_flow(dest_ids[f], (cr, cc)); lf = ref.fc_los(dest_ids[f], (cr, cc))
                    if ff is not None or lf is not None:
                        nav.pool_put(f, (cr, cc), ff, lf)
    nav.agents_set_work(work)
    nav.agents_tick(capi.TICK_VDES_FROM_POOL if from_pool else 0)
    got_vel = nav.agents_read_velocities(len(work))
    got_vpref, got_vdes, got_los = nav.agents_read_debug(len(work))

    def relerr(g, e):
        d = np.abs(g - e).max(axis=1)
        return d / np.maximum(np.abs(e).max(axis=1), 1e-3)
    if from_pool:
        check(f"agents[{cw}x{cw},n={n}] LOS bits from pool", (got_los == los).all(), f"mismatch={(got_los != los).sum()}")
        ev = relerr(got_vdes, vdes)
        check(f"agents[{cw}x{cw},n={n}] vdes from pool", (ev <= 1e-6).all(), f"max rel={ev.max():.3e} bitexact={(got_vdes == vdes).all()}")
    e1 = relerr(got_vpref, exp_vpref); e2 = relerr(got_vel, exp_vel)
    nb = (got_vel != exp_vel).any(axis=1).sum()
    check(f"agents[{cw}x{cw},n={n}] vpref", (e1 <= 1e-4).mean() >= 0.999, f"max rel={e1.max():.3e} frac>1e-4={(e1 > 1e-4).mean():.5f} bitexact={(got_vpref == exp_vpref).all(axis=1).mean():.4f}")
    check(f"agents[{cw}x{cw},n={n}] velocity", (e2 <= 1e-4).mean() >= 0.995, f"max rel={e2.max():.3e} frac>1e-4={(e2 > 1e-4).mean():.5f} not-bitexact={nb}/{len(work)} ref_secs={secs:.3f}")
    ref.close()

if __name__ == "__main__":
    nav = capi.Nav(0)
    which = sys.argv[1:] or ["flow", "portal", "general", "los", "agents"]
    if "flow" in which:
        test_flow_tile(nav, 0); test_flow_tile(nav, 1)
    if "portal" in which:
        test_flow_portal(nav, 0); test_flow_portal(nav, 1)
    if "general" in which:
        test_flow_general(nav)
    if "los" in which:
        test_los(nav)
    if "agents" in which:
        test_agents(nav, 1, 256, 1, 31, 0.02, 4.0, False)
        test_agents(nav, 1, 400, 2, 32, 0.05, 2.6, False)
        test_agents(nav, 3, 3000, 3, 33, 0.03, 2.6, True)
    print("launches:", nav.launch_count())
    print("FAILS:", fails)
    sys.exit(1 if fails else 0)
