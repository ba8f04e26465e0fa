#!/usr/bin/env python3
"""Generates tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref/libpfref.so, built
from /root/reference by `make -C oracle ref`) on the seeded cases of tests/cases.py.

Run in the build container (the reference checkout does not exist on the GPU box):
    python tests/golden/make_golden.py
The reference publishes no golden vectors / KATs of its own for this path, so
these files are what pins the oracle port and the CUDA path."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases
import pfref

capi, synth = cases.capi, cases.synth


def gen_flow_tile():
    out = {}
    for k, (seed, dens) in enumerate(((1, 0.0), (2, 0.25), (4, 0.4))):
        p, cost, reqs = cases.flow_tile_case(seed, dens)
        ref = pfref.RefMap(1, 1, p)
        assert (ref.cost_base() == cost).all()
        exp = np.stack([ref.flow_tile((0, 0), (int(q["tile_r"]), int(q["tile_c"]))) for q in reqs])
        out[f"cost{k}"] = cost; out[f"reqs{k}"] = reqs.view(np.uint8); out[f"exp{k}"] = exp
        ref.close()
    np.savez_compressed(os.path.join(HERE, "flow_tile.npz"), **out)


def gen_portal_los():
    cw = ch = 3
    p = cases.noise_map(cw, ch, 12, 0.15)
    ref = pfref.RefMap(cw, ch, p)
    cost, liid, ports = ref.cost_base(), ref.local_islands(), ref.portals()
    specs = cases.portal_specs(ports, liid, cw, limit=48)
    exp = np.stack([ref.flow_portal(s[0], s[1], s[5], s[6]) for s in specs])
    # in-place merge of a second target (nav.c:1998-2008)
    s0 = specs[0]
    base = ref.flow_tile(s0[0], (5, 5))
    merged = ref.flow_portal(s0[0], s0[1], s0[5], s0[6], inout=base)
    los_reqs = cases.los_case(cost, cw, ch, 22, ntargets=3)
    los_exp = cases.ref_los_batch(ref, los_reqs)
    np.savez_compressed(os.path.join(HERE, "portal_los.npz"), pathable=p, cost=cost, liid=liid, portals=ports,
                        islands=ref.islands(), reqs=cases.portal_reqs(specs).view(np.uint8), exp=exp,
                        merge_base=base, merge_exp=merged, los_reqs=los_reqs.view(np.uint8), los_exp=los_exp)
    ref.close()


def gen_agents(name, cw, n, nflocks, seed, dens, spacing):
    p, cost, a = cases.agent_case(cw, n, nflocks, seed, dens, spacing)
    ref = pfref.RefMap(cw, cw, p)
    dest_ids = []
    for f in range(nflocks):
        src = a["pos"][np.argmax(a["flock_of"] == f)]
        tgt = a["flock_target"][f]
        ok, did = ref.request_path((float(src[0]), float(src[1])), (float(tgt[0]), float(tgt[1])))
        dest_ids.append(did if ok else ref.dest_id((float(tgt[0]), float(tgt[1]))))
    ref.agents_set(a["pos"], a["prev_pos"], a["vel"], a["radius"], a["max_speed"], a["state"], a["flags"],
                   a["flock_of"], a["flock_target"], np.array(dest_ids, np.uint32), hz=20)
    work = np.nonzero((a["state"] != 2) & (a["state"] != 4))[0].astype(np.uint32)
    vdes = np.zeros((len(work), 2), np.float32); los = np.zeros(len(work), np.uint8)
    for _pass in range(2):       # pass 0 settles the reference's field cache (see tests/tools/gpu_check.py)
        for f in range(nflocks):
            sel = np.nonzero(a["flock_of"][work] == f)[0]
            if len(sel) == 0:
                continue
            v, l = ref.desired_velocity(dest_ids[f], a["pos"][work[sel]], a["prev_pos"][work[sel]], a["flock_target"][f])
            vdes[sel] = v; los[sel] = l
    ref.work_set(work, vdes, los, a["speed"][work])
    vel, _ = ref.velocity_work(1)
    vpref = ref.vpref()
    rng = np.random.default_rng(seed)
    qi = rng.integers(0, n, 12)
    q10 = [ref.ents_in_circle(float(a["pos"][i, 0]), float(a["pos"][i, 1]), 10.0, 512) for i in qi]
    q30 = [ref.ents_in_circle(float(a["pos"][i, 0]), float(a["pos"][i, 1]), 30.0, 128) for i in qi]
    pool_chunks, pool_flow, pool_los = [], [], []
    for f in range(nflocks):
        for cr in range(cw):
            for cc in range(cw):
                ff, _ = ref.fc_flow(dest_ids[f], (cr, cc)); lf = ref.fc_los(dest_ids[f], (cr, cc))
                if ff is None and lf is None:
                    continue
                pool_chunks.append((f, cr, cc, ff is not None, lf is not None))
                pool_flow.append(ff if ff is not None else np.zeros((64, 64), np.uint8))
                pool_los.append(lf if lf is not None else np.zeros((64, 64), np.uint8))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), pathable=p, cost=cost,
                        **{"a_" + k: v for k, v in a.items() if isinstance(v, np.ndarray)},
                        work=work, vdes=vdes, los=los, vel=vel, vpref=vpref, qi=qi,
                        q10=np.concatenate(q10), q10_len=np.array([len(x) for x in q10]),
                        q30=np.concatenate(q30), q30_len=np.array([len(x) for x in q30]),
                        pool_chunks=np.array(pool_chunks, np.int32), pool_flow=np.array(pool_flow, np.uint8),
                        pool_los=np.array(pool_los, np.uint8), liid=ref.local_islands())
    ref.close()


def gen_route():
    """N_RequestPath on a cold field cache for seeded (src, dst) pairs: the (chunk -> ff_id) mapping, the
    cached flow and LOS fields. Includes pairs whose chained LOS exercises the neighbour-snapshot
    order of field_neighbours_grid_los (field.c:2205)."""
    cw = ch = 3
    out = {}
    for k, (seed, dens) in enumerate(((81, 0.08), (5, 0.3))):
        p = cases.noise_map(cw, ch, seed, dens)
        ref = pfref.RefMap(cw, ch, p)
        cost = ref.cost_base()
        pairs = cases.route_pairs(cost, cw, ch, seed, 24)
        oks, dids, ffids, flows, loss, has = [], [], [], [], [], []
        for src, dst in pairs:
            ref.fc_clear()
            ok, did = ref.request_path(src, dst)
            oks.append(ok); dids.append(did)
            fid = np.zeros(cw * ch, np.uint64); hs = np.zeros(cw * ch, np.uint8)
            fl = np.zeros((cw * ch, 64, 64), np.uint8); ls = np.zeros((cw * ch, 64, 64), np.uint8)
            if ok:
                for c in range(cw * ch):
                    f, i = ref.fc_flow(did, (c // cw, c % cw)); l = ref.fc_los(did, (c // cw, c % cw))
                    if f is not None:
                        fl[c] = f; fid[c] = i; hs[c] |= 1
                    if l is not None:
                        ls[c] = l; hs[c] |= 2
            ffids.append(fid); flows.append(fl); loss.append(ls); has.append(hs)
        out.update({f"pathable{k}": p, f"cost{k}": cost, f"liid{k}": ref.local_islands(), f"islands{k}": ref.islands(),
                    f"pairs{k}": np.array(pairs, np.float32), f"ok{k}": np.array(oks), f"did{k}": np.array(dids, np.uint32),
                    f"ffid{k}": np.array(ffids), f"flow{k}": np.array(flows), f"los{k}": np.array(loss), f"has{k}": np.array(has)})
        ports = ref.portals()
        edges = []
        for row in ports:
            e = ref.portal_edges(0, int(row[0]) * cw + int(row[1]), int(row[2]))
            edges.append(np.concatenate([[len(e)], e.ravel()]).astype(np.uint32))
        out[f"portals{k}"] = ports; out[f"edges{k}"] = np.concatenate(edges)
        ref.close()
    np.savez_compressed(os.path.join(HERE, "route_3x3.npz"), **out)


def gen_tiles():
    """N_NewCtxForMapData from explicit tile attributes (nav.c:2284): cost_base of a ground, water
    and air layer, then the structures derived from it on the ground layer."""
    out = {}
    for k, (cw, ch, seed, terrain) in enumerate(((2, 2, 41, False), (3, 2, 42, True))):
        t = cases.tile_attr_case(cw, ch, seed, terrain)
        ref = pfref.RefMap(cw, ch, tiles=t)
        out[f"tiles{k}"] = t.astype(np.int8)
        for layer in (0, 3, 4, 8):
            out[f"cost{k}_{layer}"] = ref.cost_base(layer)
        out[f"liid{k}"] = ref.local_islands(0)
        out[f"portals{k}"] = ref.portals(0)
        out[f"islands{k}"] = ref.islands(0)
        ref.close()
    np.savez_compressed(os.path.join(HERE, "tiles.npz"), **out)


def gen_update(name, seed, hz):
    """entity_compute_update (movement.c:2303) + the movestate part of entity_apply_update (:2693)"""
    cw = 3
    p, cost, a, ms = cases.update_case(seed, hz)
    nflocks = len(a["flock_target"])
    ref = pfref.RefMap(cw, cw, p)
    dest_ids = []
    for f in range(nflocks):
        src = a["pos"][np.argmax(a["flock_of"] == f)]
        tgt = a["flock_target"][f]
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
            if len(sel) == 0:
                continue
            v, l = ref.desired_velocity(dest_ids[f], a["pos"][work[sel]], a["prev_pos"][work[sel]], a["flock_target"][f])
            vdes[sel] = v; los[sel] = l
    rng = np.random.default_rng(seed)
    vdes[rng.random(len(work)) < 0.03] = 0.0          # "navigation cannot guide the entity any closer" -> WAITING
    ref.work_set(work, vdes, los, a["speed"][work])
    vel, _ = ref.velocity_work(1)
    oi, of = ref.compute_updates(vel)
    hist, hidx = ref.apply_velocity_patch(of[:, 0:2], oi[:, 0])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), pathable=p, cost=cost,
                        **{"a_" + k: v for k, v in a.items() if isinstance(v, np.ndarray)},
                        ms=ms.view(np.uint8), work=work, vdes=vdes, los=los, vel=vel, patch_i=oi, patch_f=of,
                        hist=hist, hidx=hidx, hz=np.int32(hz))
    print(name, "flags histogram", {int(k): int(v) for k, v in zip(*np.unique(oi[:, 0], return_counts=True))},
          "states", {int(k): int(v) for k, v in zip(*np.unique(oi[:, 1], return_counts=True))})
    ref.close()


def gen_repair():
    """N_FlowFieldUpdateToNearestPathable / N_FlowFieldUpdateIslandToNearest (field.c:2247, 2307) on a map whose
    local islands were cut by blockers"""
    cw = ch = 2
    p = cases.noise_map(cw, ch, 77, 0.2)
    ref = pfref.RefMap(cw, ch, p)
    rng = np.random.default_rng(5)
    for _ in range(40):
        ref.blockers_incref(float(-rng.uniform(10, cw * 256 - 10)), float(rng.uniform(10, ch * 256 - 10)),
                            float(rng.uniform(3, 14)), 0, 0)
    ref.update()
    T, K, A, B, E = cases.repair_case(ref, cw, ch, 9)
    np.savez_compressed(os.path.join(HERE, "repair.npz"), pathable=p, cost=ref.cost_base(), blk=ref.blockers(),
                        liid=ref.local_islands(), islands=ref.islands(), targets=T.view(np.uint8), kinds=K, args=A,
                        base=B, exp=E)
    print("repair cases", len(K), "changed tiles per case", float((B != E).reshape(len(K), -1).sum(1).mean()))
    ref.close()


def gen_repair_pool():
    """N_DesiredPointSeekVelocity with its on-miss chain (nav.c:3468-3554) for entities standing on blocked tiles,
    on cut-off local islands and in chunks the first path request never touched. Steady state: the reference's
    answers on a second pass over the same positions (the first pass mutates its field cache entity by entity)."""
    cw = ch = 2
    p = cases.noise_map(cw, ch, 83, 0.15)
    ref = pfref.RefMap(cw, ch, p)
    rng = np.random.default_rng(11)
    for _ in range(45):
        ref.blockers_incref(float(-rng.uniform(10, cw * 256 - 10)), float(rng.uniform(10, ch * 256 - 10)),
                            float(rng.uniform(3, 12)), 0, 0)
    ref.update()
    cost, blk, liid = ref.cost_base(), ref.blockers(), ref.local_islands()
    img_c = synth.blocked_to_image(cost, cw, ch); img_b = synth.blocked_to_image(blk.astype(np.uint16), cw, ch) if False else \
        blk.reshape(ch, cw, 64, 64).transpose(0, 2, 1, 3).reshape(ch * 64, cw * 64)
    free = np.argwhere((img_c != 255) & (img_b == 0))
    blocked = np.argwhere((img_c != 255) & (img_b > 0))
    wall = np.argwhere(img_c == 255)
    def centre(t): return np.stack([-(t[:, 1] + 0.5) * 4.0, (t[:, 0] + 0.5) * 4.0], axis=1).astype(np.float32)
    tgt_tile = free[(free[:, 0] >= 64) & (free[:, 1] >= 64)][7]
    target = centre(tgt_tile[None])[0]
    pos = np.concatenate([centre(free[rng.integers(0, len(free), 160)]) + rng.uniform(-1.5, 1.5, (160, 2)).astype(np.float32),
                          centre(blocked[rng.integers(0, len(blocked), 100)]),
                          centre(wall[rng.integers(0, len(wall), 40)])]).astype(np.float32)
    ok, did = ref.request_path((float(pos[0, 0]), float(pos[0, 1])), (float(target[0]), float(target[1])))
    assert ok
    v1, l1 = ref.desired_velocity(did, pos, pos, target)
    v2, l2 = ref.desired_velocity(did, pos, pos, target)
    v3, l3 = ref.desired_velocity(did, pos, pos, target)
    assert (v2 == v3).all()
    print("repair_pool: pass1 != pass2 for", int((v1 != v2).any(axis=1).sum()), "agents; zero vdes", int((np.abs(v2).sum(1) == 0).sum()))
    np.savez_compressed(os.path.join(HERE, "repair_pool.npz"), pathable=p, cost=cost, blk=blk, liid=liid, pos=pos,
                        target=target, vdes=v2, los=l2, did=np.uint32(did))
    ref.close()


def gen_faction():
    """"attacking" requests (N_RequestPathAttacking, nav.c:3393): flow (tile + portal targets) and LOS fields whose
    passability follows field_tile_passable_no_enemies (field.c:179) over per-faction blocker refcounts"""
    cw = ch = 2
    p = cases.noise_map(cw, ch, 55, 0.12)
    ref = pfref.RefMap(cw, ch, p)
    rng = np.random.default_rng(3)
    wars = [(0, 1), (0, 2), (3, 1)]
    for a, b in wars:
        ref.set_war(a, b)
    blockers = []
    for _ in range(60):
        b = (float(-rng.uniform(10, cw * 256 - 10)), float(rng.uniform(10, ch * 256 - 10)), float(rng.uniform(2, 10)), int(rng.integers(0, 4)))
        blockers.append(b)
        ref.blockers_incref(b[0], b[1], b[2], b[3], 0)
    ref.update()
    cost, blk, liid, fac = ref.cost_base(), ref.blockers(), ref.local_islands(), ref.factions()
    ports = ref.portals()
    specs = cases.portal_specs(ports, liid, cw, limit=24)
    treq, texp, preq, pexp, lreq, lexp = [], [], [], [], [], []
    for f in (0, 1, 3):
        for chunk in range(cw * ch):
            cr, cc = chunk // cw, chunk % cw
            npass = np.argwhere(cost[chunk] != 255)
            for t in npass[rng.integers(0, len(npass), 4)]:
                q = capi.tile_req((cr, cc), (int(t[0]), int(t[1]))); q["faction_id"] = f
                treq.append(q); texp.append(ref.flow_tile((cr, cc), (int(t[0]), int(t[1])), faction=f))
                td = (cr, cc, int(t[0]), int(t[1]))
                i0 = len(lreq)
                ql = capi.los_req((cr, cc), td); ql["faction_id"] = f
                lreq.append(ql); lexp.append(ref.los((cr, cc), td, faction=f))
                nb = (cr, cc + 1) if cc + 1 < cw else (cr, cc - 1)
                qn = capi.los_req(nb, td, prev_index=i0, prev_chunk=(cr, cc)); qn["faction_id"] = f
                lreq.append(qn); lexp.append(ref.los(nb, td, prev=lexp[i0], prev_chunk=(cr, cc), faction=f))
        for s_ in specs:
            q = cases.portal_reqs([s_]); q["faction_id"] = f
            preq.append(q); pexp.append(ref.flow_portal(s_[0], s_[1], s_[5], s_[6], faction=f))
    enemies = np.zeros(16, np.uint16)
    for a, b in wars:
        enemies[a] |= 1 << b; enemies[b] |= 1 << a
    np.savez_compressed(os.path.join(HERE, "faction.npz"), pathable=p, cost=cost, blk=blk, liid=liid, factions=fac,
                        enemies=enemies, blockers=np.array(blockers, np.float32),
                        treq=np.concatenate(treq).view(np.uint8), texp=np.stack(texp),
                        preq=np.concatenate(preq).view(np.uint8), pexp=np.stack(pexp),
                        lreq=np.concatenate(lreq).view(np.uint8), lexp=np.stack(lexp))
    plain = np.stack([ref.flow_tile((int(q["chunk_r"][0]), int(q["chunk_c"][0])), (int(q["tile_r"][0]), int(q["tile_c"][0]))) for q in treq])
    print("faction: tile fields differing from the plain rule:", int((plain != np.stack(texp)).reshape(len(treq), -1).any(axis=1).sum()), "of", len(treq))
    ref.close()


def pack_region_reqs(reqs):
    """-> int32[n, 8] = center r c, target r c, enemies, start r c (-1 = none), noverlay; int32[sum, 2] overlay tiles"""
    rec = np.full((len(reqs), 8), -1, np.int32); ovs = []
    for i, q in enumerate(reqs):
        ov = np.zeros((0, 2), np.int32) if q["overlay"] is None else np.asarray(q["overlay"], np.int32).reshape(-1, 2)
        rec[i, 0:2] = q["center"]; rec[i, 2:4] = q["target"]; rec[i, 4] = q["enemies"]
        if q["start"] is not None:
            rec[i, 5:7] = q["start"]
        rec[i, 7] = len(ov); ovs.append(ov)
    return rec, np.concatenate(ovs) if ovs else np.zeros((0, 2), np.int32)


def gen_region():
    """Region fields: N_CellArrivalFieldCreate [+ N_CellArrivalFieldUpdateToNearestPathable]
    (field.c:2445, 2603) at the formation size 96 and at 32, and N_GroupArrivalFieldCreate (field.c:2525)"""
    cw = ch = 3
    out = {}
    for dim in (96, 32):
        p, blockers, wars, reqs = cases.region_case(21, cw, ch, 48, dim)
        ref = pfref.RefMap(cw, ch, p)
        for a, b in wars:
            ref.set_war(a, b)
        for b in blockers:
            ref.blockers_incref(b[0], b[1], b[2], b[3], 0)
        ref.update()
        cost, blk, fac = ref.cost_base(), ref.blockers(), ref.factions()
        cases.region_pick_starts(reqs, cost, blk, cw, ch, 21, dim)
        exp = np.stack([ref.cell_arrival_field(dim, q["target"], q["center"], q["enemies"], q["overlay"], q["start"]) for q in reqs])
        create_only = np.stack([ref.cell_arrival_field(dim, q["target"], q["center"], q["enemies"], q["overlay"], None) for q in reqs])
        rec, ov = pack_region_reqs(reqs)
        rng = np.random.default_rng(dim)
        gt, gc, ge, gx = [], [], [], []
        for k in range(10):
            c = np.array([-rng.uniform(-4, cw * 256 + 4), rng.uniform(-4, ch * 256 + 4)], np.float32)     # sometimes off the map
            t = np.stack([c[0] + rng.uniform(-220, 220, 16), c[1] + rng.uniform(-220, 220, 16)], 1).astype(np.float32)
            e = [0, 0b0110][k % 2]
            gt.append(t); gc.append(c); ge.append(e); gx.append(ref.group_arrival_field(dim, t, c, e))
        nfix = sum(q["start"] is not None for q in reqs)
        print("region dim %d: %d requests, %d with fix-up (changed %d), group fields %d (zero: %d)" % (
            dim, len(reqs), nfix, int((exp != create_only).reshape(len(reqs), -1).any(axis=1).sum()), len(gx),
            sum(int(not x.any()) for x in gx)))
        if dim == 96:
            out.update(pathable=p, cost=cost, blk=blk, factions=fac, blockers=np.array(blockers, np.float32))
            # TARGET_ZONE chunk fields (field_update_zone, field.c:1810) for every chunk of the map, and the consumer
            # N_DesiredGroupArrivalVelocity (nav.c:3561) after N_RequestAsyncGroupArrivalField's chunk selection
            zr = np.random.default_rng(5)
            zc = np.stack([zr.integers(0, ch * 64, 12), zr.integers(0, cw * 64, 12)], 1).astype(np.int32)
            zrad = np.array([0, 1, 2, 3, 5, 8, 12, 17, 23, 30, 45, 70], np.int32)
            zexp = np.stack([np.stack([ref.flow_field_zone((c // cw, c % cw), zc[k], int(zrad[k])) for c in range(cw * ch)])
                             for k in range(len(zc))])
            gv = []
            for k in range(6):
                cxz = np.array([-zr.uniform(-3, cw * 256 + 3), zr.uniform(-3, ch * 256 + 3)], np.float32)
                rad = int(zr.integers(2, 30))
                pos = np.stack([cxz[0] + zr.uniform(-260, 260, 400), cxz[1] + zr.uniform(-260, 260, 400)], 1).astype(np.float32)
                v, f, nb = ref.group_arrival_velocity(cxz, rad, pos)
                gv.append((cxz, rad, pos, v, f, nb))
                print("zone consumer %d: %d chunk fields, %d ok, %d at slot" % (k, nb, int((f & 1).sum()), int((f >> 1).sum())))
            print("zone fields: non-empty %d of %d" % (int(zexp.reshape(-1, 4096).any(axis=1).sum()), zexp.shape[0] * zexp.shape[1]))
            out.update(zc=zc, zrad=zrad, zexp=zexp, gv_centre=np.stack([g[0] for g in gv]), gv_radius=np.array([g[1] for g in gv], np.int32),
                       gv_pos=np.stack([g[2] for g in gv]), gv_vel=np.stack([g[3] for g in gv]), gv_flags=np.stack([g[4] for g in gv]),
                       gv_nfields=np.array([g[5] for g in gv], np.int32))
        out.update({"req%d" % dim: rec, "ov%d" % dim: ov, "exp%d" % dim: exp, "create%d" % dim: create_only,
                    "gt%d" % dim: np.stack(gt), "gc%d" % dim: np.stack(gc), "ge%d" % dim: np.array(ge, np.int32), "gx%d" % dim: np.stack(gx)})
        ref.close()
    np.savez_compressed(os.path.join(HERE, "region.npz"), **out)



def gen_demo_map():
    """The engine's own demo map (assets/maps/demo.pfmap, 4 x 4 chunks, every tile type, heights -3..9) through the
    reference's nav build: per-layer cost grids, local islands, portals and a few path requests. The tile attributes
    are read with the package's PFMAP parser (pfnav_pfmap_parse == m_al_parse_tile's fixed positions) and stored, so
    the GPU box can rebuild the PFMAP text without the asset."""
    text = open("/root/reference/assets/maps/demo.pfmap", "rb").read()
    tiles = capi.pfmap_parse(text)
    ch, cw = tiles.shape[0] // 32, tiles.shape[1] // 32
    ref = pfref.RefMap(cw, ch, tiles=tiles)
    out = dict(tiles=tiles.astype(np.int8))
    for layer in (0, 1, 3, 4, 8):
        out["cost_%d" % layer] = ref.cost_base(layer)
    out["liid_0"] = ref.local_islands(0)
    out["portals_0"] = ref.portals(0)
    out["islands_0"] = ref.islands(0)
    cost = out["cost_0"]
    pairs = cases.route_pairs(cost, cw, ch, 9, 16)
    n = cw * ch
    oks, dids, ffids, flows, loss, has = [], [], [], [], [], []
    for src, dst in pairs:
        ref.fc_clear()
        ok, did = ref.request_path(src, dst)
        oks.append(ok); dids.append(did)
        fid = np.zeros(n, np.uint64); hs = np.zeros(n, np.uint8)
        fl = np.zeros((n, 64, 64), np.uint8); ls = np.zeros((n, 64, 64), np.uint8)
        if ok:
            for c in range(n):
                f, k = ref.fc_flow(did, (c // cw, c % cw)); l = ref.fc_los(did, (c // cw, c % cw))
                if f is not None:
                    fl[c] = f; fid[c] = k; hs[c] |= 1
                if l is not None:
                    ls[c] = l; hs[c] |= 2
        ffids.append(fid); flows.append(fl); loss.append(ls); has.append(hs)
    out.update(pairs=np.array(pairs, np.float32), ok=np.array(oks), did=np.array(dids, np.uint32), ffid=np.array(ffids),
               flow=np.array(flows), los=np.array(loss), has=np.array(has))
    print("demo map: %dx%d chunks, impassable ground tiles %d, portals %d, routes ok %d of %d" % (
        ch, cw, int((cost == 255).sum()), len(out["portals_0"]), int(sum(oks)), len(oks)))
    np.savez_compressed(os.path.join(HERE, "demo_map.npz"), **out)
    ref.close()



def gen_targets():
    """TARGET_ENTITY / TARGET_ENEMIES chunk fields (field_update_entity field.c:1609, field_update_enemies :1540) on
    reference layers 0 (1x1) and 2 (5x5: contour rings), through a nav_unit_query_ctx over 120 entities"""
    cw = ch = 3
    p, blockers, wars, _ = cases.region_case(7, cw, ch, 4, 96)
    ref = pfref.RefMap(cw, ch, p)
    for a, b in wars:
        ref.set_war(a, b)
    for b in blockers[:30]:
        ref.blockers_incref(b[0], b[1], b[2], b[3], 0)
    ref.update()
    t = cases.target_case(4, cw, ch)
    n = len(t["radius"])
    z = np.zeros((n, 2), np.float32)
    ref.agents_set(t["pos"], t["pos"], z, t["radius"], np.ones(n, np.float32), np.zeros(n, np.int32), t["flags"],
                   np.full(n, -1, np.int32), np.zeros((0, 2), np.float32), np.zeros(0, np.uint32))
    ref.agents_set_factions(t["factions"])
    uids = np.arange(0, n, 15)
    out = dict(pathable=p, blockers=np.array(blockers[:30], np.float32), wars=np.array(wars, np.int32), uids=uids, **t)
    for L in (0, 2):
        out["cost_%d" % L] = ref.cost_base(L); out["blk_%d" % L] = ref.blockers(L)
        out["ent_%d" % L] = np.stack([np.stack([ref.flow_field_entity((c // cw, c % cw), int(u), layer=L) for c in range(cw * ch)]) for u in uids])
        out["foe_%d" % L] = np.stack([np.stack([ref.flow_field_enemies((c // cw, c % cw), f, layer=L) for c in range(cw * ch)]) for f in range(4)])
        print("targets layer %d: entity fields non-empty %d of %d, enemies %d of %d" % (
            L, int(out["ent_%d" % L].reshape(-1, 4096).any(axis=1).sum()), len(uids) * cw * ch,
            int(out["foe_%d" % L].reshape(-1, 4096).any(axis=1).sum()), 4 * cw * ch))
    np.savez_compressed(os.path.join(HERE, "targets.npz"), **out)
    ref.close()



def gen_stress():
    """The reference's own stress scenario (scripts/test_stress.py): assets/maps/plain.pfmap centred at the origin
    (M_CenterAtOrigin, map.c:420: pos = (+512, 0, -512)), 2 x 256 units marching at each other. One movement tick:
    both path requests (flow + LOS fields of the field cache), vdes / LOS per unit, the velocity pass. Units start with
    half their speed towards the goal (+ seeded jitter) so that the obstacle avoidance has moving neighbours."""
    text = open("/root/reference/assets/maps/plain.pfmap", "rb").read()
    tiles = capi.pfmap_parse(text)
    ch, cw = tiles.shape[0] // 32, tiles.shape[1] // 32
    mx, mz = cw * 256 / 2.0, -(ch * 256 / 2.0)
    ref = pfref.RefMap(cw, ch, tiles=tiles, map_x=mx, map_z=mz)
    pos, radius, flock_of, targets = cases.stress_layout()
    n = len(pos)
    rng = np.random.default_rng(12)
    d = targets[flock_of] - pos
    vel = (d / np.maximum(np.sqrt((d * d).sum(1, keepdims=True)), 1e-6) * (0.5 * 20.0 / 20)).astype(np.float32)
    vel = (vel + (rng.random((n, 2)).astype(np.float32) - 0.5) * np.float32(0.05)).astype(np.float32)
    a = dict(pos=pos, prev_pos=(pos - vel).astype(np.float32), vel=vel, radius=radius, max_speed=np.full(n, 20.0, np.float32),
             speed=np.full(n, 20.0, np.float32), state=np.zeros(n, np.int32), flags=np.full(n, 1 << 3, np.uint32),
             flock_of=flock_of, flock_target=targets)
    nc = cw * ch
    dest_ids, oks, ffids, flows, loss, has, srcs = [], [], [], [], [], [], []
    for f in range(2):
        src = pos[np.argmax(flock_of == f)]
        ref.fc_clear()
        ok, did = ref.request_path((float(src[0]), float(src[1])), (float(targets[f][0]), float(targets[f][1])))
        assert ok
        fid = np.zeros(nc, np.uint64); hs = np.zeros(nc, np.uint8)
        fl = np.zeros((nc, 64, 64), np.uint8); ls = np.zeros((nc, 64, 64), np.uint8)
        for c in range(nc):
            ff, k = ref.fc_flow(did, (c // cw, c % cw)); lf = ref.fc_los(did, (c // cw, c % cw))
            if ff is not None:
                fl[c] = ff; fid[c] = k; hs[c] |= 1
            if lf is not None:
                ls[c] = lf; hs[c] |= 2
        dest_ids.append(did); oks.append(ok); ffids.append(fid); flows.append(fl); loss.append(ls); has.append(hs); srcs.append(src)
    ref.fc_clear()
    for f in range(2):
        ref.request_path((float(srcs[f][0]), float(srcs[f][1])), (float(targets[f][0]), float(targets[f][1])))
    ref.agents_set(a["pos"], a["prev_pos"], a["vel"], a["radius"], a["max_speed"], a["state"], a["flags"],
                   a["flock_of"], a["flock_target"], np.array(dest_ids, np.uint32), hz=20)
    work = np.arange(n, dtype=np.uint32)
    vdes = np.zeros((n, 2), np.float32); los = np.zeros(n, np.uint8)
    for _pass in range(2):
        for f in range(2):
            sel = np.nonzero(flock_of == f)[0]
            v, l = ref.desired_velocity(dest_ids[f], pos[sel], a["prev_pos"][sel], targets[f])
            vdes[sel] = v; los[sel] = l
    # the settled field cache (what the second pass read)
    pool_chunks, pool_flow, pool_los = [], [], []
    for f in range(2):
        for c in range(nc):
            ff, _ = ref.fc_flow(dest_ids[f], (c // cw, c % cw)); lf = ref.fc_los(dest_ids[f], (c // cw, c % cw))
            if ff is None and lf is None:
                continue
            pool_chunks.append((f, c // cw, c % cw, ff is not None, lf is not None))
            pool_flow.append(ff if ff is not None else np.zeros((64, 64), np.uint8))
            pool_los.append(lf if lf is not None else np.zeros((64, 64), np.uint8))
    ref.work_set(work, vdes, los, a["speed"][work])
    vel_out, _ = ref.velocity_work(1)
    vpref = ref.vpref()
    print("stress: %d units, map origin (%.0f, %.0f), units with LOS %d, zero vdes %d, zero new velocity %d, pool fields %d" % (
        n, mx, mz, int(los.sum()), int((np.abs(vdes).sum(1) == 0).sum()), int((np.abs(vel_out).sum(1) == 0).sum()), len(pool_chunks)))
    np.savez_compressed(os.path.join(HERE, "stress.npz"), tiles=tiles.astype(np.int8), cost=ref.cost_base(), liid=ref.local_islands(),
                        islands=ref.islands(), map_origin=np.array([mx, mz], np.float32),
                        **{"a_" + k: v for k, v in a.items()}, work=work, vdes=vdes, los=los, vel=vel_out, vpref=vpref,
                        pairs=np.array([[srcs[f], targets[f]] for f in range(2)], np.float32), ok=np.array(oks), did=np.array(dest_ids, np.uint32),
                        ffid=np.array(ffids), flow=np.array(flows), los_f=np.array(loss), has=np.array(has),
                        pool_chunks=np.array(pool_chunks, np.int32), pool_flow=np.array(pool_flow, np.uint8), pool_los=np.array(pool_los, np.uint8))
    ref.close()



if __name__ == "__main__":
    gen_stress()
    gen_targets()
    gen_demo_map()
    gen_region()
    gen_faction()
    gen_repair_pool()
    gen_repair()
    gen_update("update_hz20", 61, 20)
    gen_update("update_hz10", 62, 10)
    gen_tiles()
    gen_route()
    gen_flow_tile()
    gen_portal_los()
    gen_agents("agents_1x1", 1, 256, 1, 31, 0.02, 4.0)
    gen_agents("agents_dense", 1, 400, 2, 32, 0.05, 2.6)
    gen_agents("agents_3x3", 3, 1500, 3, 33, 0.03, 2.6)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
