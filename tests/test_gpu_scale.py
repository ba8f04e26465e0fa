"""GPU suite at BASELINE.json sizes against the COMPILED REFERENCE (oracle/_ref/libpfref.so travels to the GPU box):
seeded >= 10k-agent samples of the bench's own C2 / C3 populations with the whole population present as neighbours,
desired velocities out of the field pool incl. the on-miss chain on the 1024 x 1024 map, 1 k dynamic obstacles
(config C5's churn) and multi-tick device-resident trajectories. Every comparison is an asserted maximum: integers
bit-exact, velocities within north_star's 1e-4 relative for EVERY agent."""
import importlib
import os
import sys

import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu
capi, synth = cases.capi, cases.synth
VEL_RTOL = 1e-4
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


@pytest.fixture(scope="module")
def bench():
    return importlib.import_module("bench")


@pytest.fixture(scope="module")
def c2map(pf, pfref, bench):
    """the bench's 16 x 16-chunk map in the reference (N_NewCtxForMapData: ~30 s, 2 GB) -- shared by the tests below"""
    bench.set_workload("C2")
    p = synth.make_map(bench.CHUNKS, bench.CHUNKS, bench.MAP_SEED)
    ref = pfref.RefMap(bench.CHUNKS, bench.CHUNKS, p)
    yield p, ref
    ref.close()


def _population(pf, bench, workload):
    bench.set_workload(workload)
    W = bench.build_workload(pf, 1, 0)
    bench.set_workload("C2")
    return W


@pytest.mark.parametrize("workload", ["C2", "C3"])
def test_bench_population_sample_vs_reference(pf, nav, c2map, bench, workload):
    """10 000 seeded agents of the bench population (100 k agents k10 ~ 21 / 1 M agents k10 ~ 39; every other agent is
    present as a neighbour): position-index query order (G_Pos_EntsInCircleFrom), preferred velocity and new velocity
    of move_velocity_work (movement.c:3395-3466) -- single-pass and two-phase kernels."""
    p, ref = c2map
    W = _population(pf, bench, workload)
    a, n = W["agents"], W["n_total"]
    assert (ref.cost_base() == W["cost"]).all()
    rng = np.random.default_rng(7)
    nsample = 10_000
    work = np.sort(rng.choice(n, nsample, replace=False)).astype(np.uint32)
    d = a["flock_target"][a["flock_of"]] - a["pos"]
    vdes_all = (d / np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-6)).astype(np.float32)
    los_all = (rng.random(n) < 0.2).astype(np.uint8)
    ref.agents_set(a["pos"], a["prev_pos"], a["vel"], a["radius"], a["max_speed"], a["state"], a["flags"],
                   a["flock_of"], a["flock_target"], np.arange(W["nflocks"], dtype=np.uint32), hz=20)
    ref.work_set(work, vdes_all[work], los_all[work], a["speed"][work])
    evel, _ = ref.velocity_work(os.cpu_count())
    evpref = ref.vpref()
    nav.map_create(bench.CHUNKS, bench.CHUNKS, 1); nav.map_upload_layer(0, W["cost"])
    aa = dict(a); aa["vdes"] = vdes_all; aa["has_los"] = los_all.astype(np.uint32)
    rec, fl = capi.pack_agents(aa)
    nav.agents_upload(rec, fl, 20)
    nav.agents_set_work(work)
    try:
        for mode in (0, 2):
            nav.set_two_phase(mode)
            nav.agents_tick(0)
            vel = nav.agents_read_velocities(nsample)
            vpref, _, _ = nav.agents_read_debug(nsample)
            e_vp, e_v = cases.relerr(vpref, evpref), cases.relerr(vel, evel)
            assert e_vp.max() <= VEL_RTOL, (workload, mode, e_vp.max(), work[np.nonzero(e_vp > VEL_RTOL)[0][:8]])
            assert e_v.max() <= VEL_RTOL, (workload, mode, e_v.max(), work[np.nonzero(e_v > VEL_RTOL)[0][:8]])
    finally:
        nav.set_two_phase(1)
    # the sample is representative of the hard cases: capped neighbour lists and agents without an admissible velocity
    assert (np.linalg.norm(evel, axis=1) == 0).mean() > 0.02
    for i in work[:300]:
        x, z = float(a["pos"][i, 0]), float(a["pos"][i, 1])
        for r, cap in ((10.0, 512), (30.0, 128)):
            got, exp = nav.ents_in_circle(x, z, r, cap), ref.ents_in_circle(x, z, r, cap)
            assert len(got) == len(exp) and (got == exp).all(), (workload, int(i), r)


def test_windowed_cohesion_equals_member_list(pf, nav, bench):
    """cohesion_force through the position index with the exp(-0.12 d) cut-off (+ per-entity fall-back) against the plain
    sum over the flock's member list, on the bench's C2 population and on a C4-style flock of 60 000: preferred velocities
    agree to float rounding, far inside the 1e-4 budget"""
    for workload, nsel in (("C2", 20_000), ("C4", 20_000)):
        bench.set_workload(workload)
        try:
            W = bench.build_workload(pf, 1, 0) if workload == "C2" else None
            if W is None:
                # one C4 cell: 60 000 agents of radius 1.0 in one flock on an 8 x 8-chunk map
                p = synth.make_map(8, 8, 0x5EED0004)
                cost = synth.cost_from_pathable(p, 8, 8)
                a = synth.make_agents(cost, 8, 8, 60_000, 1, 0x5EED0004, radius=1.0, spacing=3.9)
                W = dict(cost=cost, agents=a, n_total=60_000, chunks=8)
            else:
                W["chunks"] = bench.CHUNKS
        finally:
            bench.set_workload("C2")
        a, n, cw = W["agents"], W["n_total"], W["chunks"]
        d = a["flock_target"][a["flock_of"]] - a["pos"]
        aa = dict(a); aa["vdes"] = (d / np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-6)).astype(np.float32)
        rec, fl = capi.pack_agents(aa)
        nav.map_create(cw, cw, 1); nav.map_upload_layer(0, W["cost"])
        nav.agents_upload(rec, fl, 20)
        work = np.sort(np.random.default_rng(3).choice(n, nsel, replace=False)).astype(np.uint32)
        nav.agents_set_work(work)
        out = {}
        try:
            for mode in (2, 1):
                nav.set_cohesion_mode(mode)
                nav.agents_tick(0)
                out[mode] = (nav.agents_read_debug(nsel)[0], nav.agents_read_velocities(nsel))
        finally:
            nav.set_cohesion_mode(0)
        e_vp = cases.relerr(out[1][0], out[2][0]); e_v = cases.relerr(out[1][1], out[2][1])
        assert e_vp.max() <= 2e-5, (workload, e_vp.max())
        assert e_v.max() <= VEL_RTOL, (workload, e_v.max())


def test_c2_desired_velocity_from_pool_vs_reference(pf, nav, c2map, bench):
    """A2 at full size: N_RequestPath for four flocks of the C2 population on the 1024 x 1024 map (routes of up to ~20
    hops), then N_HasDestLOS + N_DesiredPointSeekVelocity for 2 000 of their agents incl. the on-miss chain
    (compute_los_state / compute_desired_velocity, movement.c:4129-4180) -- bit-exact -- and the velocity pass on top."""
    p, ref = c2map
    W = _population(pf, bench, "C2")
    a, n = W["agents"], W["n_total"]
    flocks = [0, 5, 9, 14]
    rng = np.random.default_rng(11)
    work = np.sort(np.concatenate([rng.choice(np.nonzero(a["flock_of"] == f)[0], 500, replace=False) for f in flocks])).astype(np.uint32)
    dest_ids = np.zeros(W["nflocks"], np.uint32)
    nav.map_create(bench.CHUNKS, bench.CHUNKS, 1); nav.map_upload_layer(0, W["cost"]); nav.map_build_nav(0); nav.route_build(0)
    nav.pool_create(len(flocks), 2048)
    dest_index = np.full(W["nflocks"], -1, np.int32)
    for k, f in enumerate(flocks):
        src = a["pos"][a["flock_of"] == f].mean(axis=0)
        src = a["pos"][a["flock_of"] == f][np.argmin(np.linalg.norm(a["pos"][a["flock_of"] == f] - src, axis=1))]
        dst = a["flock_target"][f]
        ok_r, did_r = ref.request_path((float(src[0]), float(src[1])), (float(dst[0]), float(dst[1])))
        ok, did, nf, nl = nav.pool_request_path(k, (float(src[0]), float(src[1])), (float(dst[0]), float(dst[1])))
        assert ok == ok_r and (not ok or did == did_r), f
        dest_ids[f] = did_r
        dest_index[f] = k
        for c in range(bench.CHUNKS ** 2):            # the route's field set equals the reference's cache, byte for byte
            fl_r, ffid_r = ref.fc_flow(did_r, (c // bench.CHUNKS, c % bench.CHUNKS))
            fl_g, lo_g, ffid_g = nav.pool_get(k, (c // bench.CHUNKS, c % bench.CHUNKS))
            assert (fl_r is None) == (fl_g is None), (f, c)
            if fl_r is not None:
                assert ffid_g == ffid_r and (fl_g == fl_r).all(), (f, c)
            lo_r = ref.fc_los(did_r, (c // bench.CHUNKS, c % bench.CHUNKS))
            assert (lo_r is None) == (lo_g is None) and (lo_r is None or (lo_g == lo_r).all()), (f, c)
    ref.agents_set(a["pos"], a["prev_pos"], a["vel"], a["radius"], a["max_speed"], a["state"], a["flags"],
                   a["flock_of"], a["flock_target"], dest_ids, hz=20)
    ref.work_set(work, np.zeros((len(work), 2), np.float32), np.zeros(len(work), np.uint8), a["speed"][work])
    # The reference serves misses inline, entity by entity, so within the tick that fills the cache an entity can read a
    # field before a later entity's request merges another portal into it; from the next tick on every entity reads the
    # settled fields. The comparison is against that steady state (second pass), which is what the device pool converges to.
    ref.desired_from_cache()
    evdes, elos = ref.desired_from_cache()
    evel, _ = ref.velocity_work(os.cpu_count())
    aa = dict(a); aa["flock_dest_index"] = dest_index
    rec, fl = capi.pack_agents(aa)
    nav.agents_upload(rec, fl, 20)
    nav.agents_set_work(work)
    for _ in range(8):
        nav.agents_tick(capi.TICK_VDES_FROM_POOL)
        nreq, nrep = nav.pool_repair()
        if nreq + nrep == 0:
            break
    nav.agents_tick(capi.TICK_VDES_FROM_POOL)
    vel = nav.agents_read_velocities(len(work))
    _, vdes, los = nav.agents_read_debug(len(work))
    assert (los == elos).all(), np.nonzero(los != elos)[0][:10]
    bad = np.nonzero((vdes != evdes).any(axis=1))[0]
    assert len(bad) == 0, (len(bad), work[bad[:10]], vdes[bad[:3]], evdes[bad[:3]])
    assert (np.abs(evdes).sum(axis=1) > 0).mean() > 0.95
    assert cases.relerr(vel, evel).max() <= VEL_RTOL


def test_c5_churn_vs_reference(pf, nav, c2map, bench):
    """config C5's dynamic-obstacle churn on the 1024 x 1024 map: 1 000 circular blockers placed, then every one moved
    by one tile (N_BlockersDecref + N_BlockersIncref) over three ticks with N_Update in between: blocker refcounts and
    local islands of all 256 chunks bit-exact after every commit; then flow (tile + portal) and chained LOS fields of
    dirtied chunks, and the velocity pass of 3 000 agents standing among the blockers."""
    p, ref = c2map
    W = _population(pf, bench, "C2")
    a, n = W["agents"], W["n_total"]
    cw = bench.CHUNKS
    nav.map_create(cw, cw, 1); nav.map_upload_layer(0, W["cost"]); nav.map_build_nav(0); nav.route_build(0)
    rng = np.random.default_rng(55)
    # blockers of radius 6 near the agents of four flocks (so that the tile probes of the steering see them) + anywhere
    near = a["pos"][rng.choice(np.nonzero(np.isin(a["flock_of"], [1, 4, 7, 12]))[0], 600, replace=False)] + rng.normal(scale=25.0, size=(600, 2))
    far = np.stack([-rng.uniform(20, cw * 256 - 20, 400), rng.uniform(20, cw * 256 - 20, 400)], 1)
    bpos = np.concatenate([near, far]).astype(np.float32)
    bpos[:, 0] = np.clip(bpos[:, 0], -(cw * 256 - 20), -20); bpos[:, 1] = np.clip(bpos[:, 1], 20, cw * 256 - 20)
    placed = []
    try:
        for x, z in bpos:
            ref.blockers_incref(float(x), float(z), 6.0); nav.blockers_incref(float(x), float(z), 6.0, 0, capi.FLAG_MOVABLE)
            placed.append((float(x), float(z)))
        for tick in range(3):
            ref.update(); nd = nav.map_commit()
            assert (nav.blockers(0) == ref.blockers()).all(), tick
            assert (nav.local_islands(0) == ref.local_islands()).all(), tick
            assert nd > 0
            step = rng.integers(-1, 2, size=(len(placed), 2)).astype(np.float32) * 4.0
            for k, (x, z) in enumerate(placed):
                nx = float(np.clip(x + step[k, 0], -(cw * 256 - 20), -20)); nz = float(np.clip(z + step[k, 1], 20, cw * 256 - 20))
                ref.blockers_decref(x, z, 6.0); nav.blockers_decref(x, z, 6.0, 0, capi.FLAG_MOVABLE)
                ref.blockers_incref(nx, nz, 6.0); nav.blockers_incref(nx, nz, 6.0, 0, capi.FLAG_MOVABLE)
                placed[k] = (nx, nz)
        ref.update(); nav.map_commit()
        blk, liid = ref.blockers(), ref.local_islands()
        assert (nav.blockers(0) == blk).all() and (nav.local_islands(0) == liid).all()
        # fields of chunks that hold blockers: tile targets, portal targets, chained LOS
        dirty = np.nonzero(blk.reshape(cw * cw, -1).any(axis=1))[0]
        assert len(dirty) >= 100
        ports = ref.portals()
        reqs, exp = [], []
        for c in dirty[:48]:
            free = np.argwhere((W["cost"][c] != 255) & (blk[c] == 0))
            t = free[rng.integers(len(free))]
            reqs.append(capi.tile_req((c // cw, c % cw), (int(t[0]), int(t[1]))))
            exp.append(ref.flow_tile((c // cw, c % cw), (int(t[0]), int(t[1]))))
        specs = [s for s in cases.portal_specs(ports, liid, cw) if (s[0][0] * cw + s[0][1]) in set(dirty[:24].tolist())][:64]
        for s in specs:
            reqs.append(cases.portal_reqs([s]))
            exp.append(ref.flow_portal(s[0], s[1], s[5], s[6]))
        got = nav.flow_fields_update(np.concatenate(reqs))
        assert (got == np.stack(exp)).all(), np.nonzero((got != np.stack(exp)).reshape(len(exp), -1).any(axis=1))[0][:10]
        eff = np.where(blk > 0, 255, W["cost"]).astype(np.uint8)
        lr = cases.los_case(eff, cw, cw, 77, ntargets=6)
        assert (nav.los_fields_create(lr) == cases.ref_los_batch(ref, lr)).all()
        # velocity pass among the blockers (nullify_impass_components reads the blocked tiles, movement.c:1831)
        work = np.sort(rng.choice(np.nonzero(np.isin(a["flock_of"], [1, 4, 7, 12]))[0], 3000, replace=False)).astype(np.uint32)
        d = a["flock_target"][a["flock_of"]] - a["pos"]
        vdes_all = (d / np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-6)).astype(np.float32)
        ref.agents_set(a["pos"], a["prev_pos"], a["vel"], a["radius"], a["max_speed"], a["state"], a["flags"],
                       a["flock_of"], a["flock_target"], np.arange(W["nflocks"], dtype=np.uint32), hz=20)
        ref.work_set(work, vdes_all[work], np.zeros(len(work), np.uint8), a["speed"][work])
        evel, _ = ref.velocity_work(os.cpu_count())
        evpref = ref.vpref()
        aa = dict(a); aa["vdes"] = vdes_all
        rec, fl = capi.pack_agents(aa)
        nav.agents_upload(rec, fl, 20); nav.agents_set_work(work); nav.agents_tick(0)
        vel = nav.agents_read_velocities(len(work)); vpref, _, _ = nav.agents_read_debug(len(work))
        assert cases.relerr(vpref, evpref).max() <= VEL_RTOL and cases.relerr(vel, evel).max() <= VEL_RTOL
        # the steering did see blocked tiles
        assert (np.abs(vpref - (a["vel"][work] + 0)).sum() > 0)
    finally:
        for x, z in placed:                               # leave the shared reference map as it was
            ref.blockers_decref(x, z, 6.0)
        ref.update()


def _vel_ok(vel, evel, pos):
    """north_star's 1e-4 relative -- or ONE float spacing of the entity's position, the reference's own noise floor:
    G_ClearPath_NewVelocity works in absolute map coordinates and returns `chosen point - position` (clearpath.c:694-715),
    so its result is quantised at ulp(|pos|) (3e-5 wu at |pos| = 256..512) whatever the velocity's own magnitude; a
    1e-7 relative difference in a preferred velocity (cohesion summation order) can move the chosen point by that one
    step. Returns (all within bounds, number of entities that needed the floor)."""
    d = np.abs(vel - evel).max(axis=1)
    rel_ok = d <= VEL_RTOL * np.maximum(np.abs(evel).max(axis=1), 1e-3)
    floor_ok = d <= np.spacing(np.abs(pos).max(axis=1).astype(np.float32))
    return bool((rel_ok | floor_ok).all()), int((~rel_ok & floor_ok).sum())


def _run_trajectory(nav, ref, a, ms, cw, cost, hz, nticks, with_blockers, follow=False):
    """tick -> entity_compute_update -> entity_apply_update -> next snapshot, `nticks` times, on both sides; the engine's
    share of the apply (entity_block: N_BlockersIncref for patches with next_block, then N_Update) is done by the test
    on our side. The device state is uploaded once and never touched again. follow=True: before every tick after the
    first the REFERENCE is re-seeded with the device's state (positions, velocities, movestates), so that every tick is
    compared on identical inputs; free-running, the two sides are two chaotic trajectories that separate by one
    position ulp per tick from the first ClearPath rounding step on (tests/tools/diag_traj.py prints both).
    Returns the per-tick maxima of the position / velocity deviation."""
    n = len(a["radius"])
    dest_ids = np.array([ref.dest_id((float(t[0]), float(t[1]))) for t in a["flock_target"]], np.uint32)
    ref.agents_set(a["pos"], a["prev_pos"], a["vel"], a["radius"], a["max_speed"], a["state"], a["flags"],
                   a["flock_of"], a["flock_target"], dest_ids, hz=hz)
    ref.movestate_set(ms["next_pos"][:, [0, 2]], ms["next_rot"], ms["step"], ms["left"], ms["vel_hist"], ms["vel_hist_idx"],
                      np.zeros(n, np.int32), np.zeros(n, np.int32), ms["combat_facing"])
    nav.map_create(cw, cw, 1); nav.map_upload_layer(0, cost); nav.map_build_nav(0); nav.route_build(0)
    nflocks = len(a["flock_target"])
    nav.pool_create(nflocks, nflocks * cw * cw)
    aa = dict(a); aa["flock_dest_index"] = np.arange(nflocks, dtype=np.int32)
    rec, fl = capi.pack_agents(aa)
    nav.agents_upload(rec, fl, hz)
    nav.agents_upload_movestate(ms)
    state = a["state"].copy()
    snapshot = a["pos"].copy()              # gamestate.positions of the current tick (what entity_block reads, movement.c:583)
    out = []
    # settle both field caches before the first compared tick (the reference fills its cache inline while it walks the
    # entities, see test_c2_desired_velocity_from_pool_vs_reference)
    work = np.nonzero((state != 2) & (state != 4))[0].astype(np.uint32)
    ref.work_set(work, np.zeros((len(work), 2), np.float32), np.zeros(len(work), np.uint8), a["speed"][work])
    ref.desired_from_cache()
    nfloor = 0
    for tick in range(nticks):
        work = np.nonzero((state != 2) & (state != 4))[0].astype(np.uint32)
        if len(work) == 0:
            break
        # reference
        if follow and tick:
            ref.agents_set(got["pos"], got["prev_pos"], got["velocity"], a["radius"], a["max_speed"], got["state"], a["flags"],
                           a["flock_of"], a["flock_target"], dest_ids, hz=hz)
            ref.movestate_set(gms["next_pos"][:, [0, 2]], gms["next_rot"], gms["step"], gms["left"], gms["vel_hist"],
                              gms["vel_hist_idx"], np.zeros(n, np.int32), np.zeros(n, np.int32), gms["combat_facing"])
        ref.work_set(work, np.zeros((len(work), 2), np.float32), np.zeros(len(work), np.uint8), a["speed"][work])
        evdes, elos = ref.desired_from_cache()
        evel, _ = ref.velocity_work(os.cpu_count())
        ref.update_and_apply()
        est = ref.state_get(n)
        # device
        nav.agents_set_work(work)
        for _ in range(8):
            nav.agents_tick(capi.TICK_VDES_FROM_POOL)
            nreq, nrep = nav.pool_repair()
            if nreq + nrep == 0:
                break
        nav.agents_tick(capi.TICK_VDES_FROM_POOL)
        vel = nav.agents_read_velocities(len(work))
        _, vdes, los = nav.agents_read_debug(len(work))
        assert (los == elos).all(), (tick, np.nonzero(los != elos)[0][:10])
        if tick == 0:
            assert (vdes == evdes).all(), (tick, np.nonzero((vdes != evdes).any(axis=1))[0][:10])
        else:       # positions carry the <= 1e-6 relative deviation of the previous ticks' velocities into the blend weights
            assert np.abs(vdes - evdes).max() <= VEL_RTOL, (tick, np.abs(vdes - evdes).max(), np.nonzero(np.abs(vdes - evdes).max(axis=1) > VEL_RTOL)[0][:10])
        ok, nf = _vel_ok(vel, evel, snapshot[work])
        assert ok, (tick, cases.relerr(vel, evel).max())
        nfloor += nf
        nav.agents_compute_updates()
        patches = nav.agents_read_patches(len(work))
        nav.agents_apply_updates()
        nav.agents_rebuild_index()
        got, gms = nav.agents_read_state(n)
        assert (got["state"] == est["state"]).all(), (tick, np.nonzero(got["state"] != est["state"])[0][:10])
        e_pos = np.abs(got["pos"] - est["pos"]).max(); e_prev = np.abs(got["prev_pos"] - est["prev_pos"]).max()
        e_vel = cases.relerr(got["velocity"], est["vel"]).max()
        # positions are world coordinates of magnitude <= 1e3: 1e-4 relative of the per-tick displacement (<= 1 wu)
        assert e_pos <= 1e-4 and e_prev <= 1e-4 and _vel_ok(got["velocity"], est["vel"], snapshot)[0], (tick, e_pos, e_prev, e_vel)
        out.append((len(work), float(e_pos), float(e_vel)))
        if with_blockers:
            # the engine side of entity_apply_update: entity_finish_moving -> entity_block (movement.c:685, 580)
            stopped = np.nonzero(((patches["flags"] & 1) != 0) & (patches["next_block"] != 0) &
                                 ((a["flags"][work] & capi.FLAG_GARRISONED) == 0))[0]
            for w in stopped:
                u = work[w]
                # G_Pos_GetXZFrom(gamestate.positions): the snapshot position of this tick, not the new one (movement.c:583)
                nav.blockers_incref(float(snapshot[u, 0]), float(snapshot[u, 1]), float(a["radius"][u]), 0, int(a["flags"][u]))
            nav.map_commit()
            assert (nav.blockers(0) == ref.blockers()).all(), tick
            assert (nav.local_islands(0) == ref.local_islands()).all(), tick
        state = est["state"]
        snapshot = got["pos"].copy()
    # the position-ulp floor is a rare event (a rounding step of ClearPath's absolute-coordinate result), not a blanket
    assert nfloor <= max(2, nticks // 2), nfloor
    return out


@pytest.mark.parametrize("hz", [20, 10])
def test_device_resident_trajectory_vs_reference(pf, nav, pfref, hz):
    """positions never leave the device between ticks. Eight ticks of a marching crowd (1 500 agents, three
    flocks, 3 x 3 chunks, goals far away: pure motion, heading gate, interpolation at 10 Hz) and one tick of an arriving
    crowd (arrivals, adjacent-arrived cascade, WAITING, blockers taken by the stopped entities and the N_Update that
    follows) against the reference's own entity_compute_update / entity_apply_update (movement.c:2303-2766)."""
    cw = 3
    # marching crowd
    p, cost, a = cases.agent_case(cw, 1500, 3, 3131, 0.03, 2.6)
    a["state"][:] = 0
    n = len(a["radius"])
    ms = np.zeros(n, capi.MOVESTATE)
    ms["next_pos"][:, 0] = a["pos"][:, 0]; ms["next_pos"][:, 2] = a["pos"][:, 1]
    ms["step"] = 1.0 / (20 // hz)
    ms["next_rot"] = cases.dir_quat(a["vel"] + 1e-6)
    ms["combat_facing"] = ms["next_rot"]
    ms["vel_hist"] = np.repeat(a["vel"][:, None, :], 14, axis=1)
    ref = pfref.RefMap(cw, cw, p)
    try:
        out = _run_trajectory(nav, ref, a, ms, cw, cost, hz, 8, with_blockers=True, follow=True)
        assert len(out) == 8 and out[-1][0] > 1000, out
    finally:
        ref.close()
    # arriving crowd: one tick on a settled cache -- arrivals, adjacent-arrived cascade, WAITING, then the engine's share of the
    # apply on our side (entity_block -> N_BlockersIncref for every stopped entity) and the N_Update that follows:
    # identical states, positions, blocker refcounts and local islands. (Later ticks of this scenario re-request half the
    # fields through the on-miss chain while the reference is still walking its entities; those transients depend on the
    # reference's serial order, see test_c2_desired_velocity_from_pool_vs_reference.)
    if hz == 20:
        p, cost, a, ms = cases.update_case(4242, hz)
        ref = pfref.RefMap(cw, cw, p)
        try:
            out = _run_trajectory(nav, ref, a, ms, cw, cost, hz, 1, with_blockers=True)
            assert len(out) == 1 and out[0][0] > 1000, out
        finally:
            ref.close()
