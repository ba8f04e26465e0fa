"""GPU suite, multi-GPU product path: the partition of the entities over several contexts must not
change a single bit of the result. The in-process group transport runs on ONE GPU (two or three contexts on device 0;
device 1.. are used when the box has them), so this is part of the normal `-m gpu` run; the NCCL transport (one process
per GPU) is checked by tests/tools/mgpu_nccl_check.py under torchrun (needs >= 2 GPUs)."""
import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu
capi, synth = cases.capi, cases.synth


def _devices(world):
    import torch
    n = torch.cuda.device_count()
    return [i % n for i in range(world)]


def _setup(nav, cost, cw):
    nav.map_create(cw, cw, 1); nav.map_upload_layer(0, cost); nav.map_build_nav(0); nav.route_build(0)


@pytest.mark.parametrize("world", [2, 3])
def test_group_equals_single_context(pf, world):
    """N contexts, each holding only its own range of the population + one all-gather of the 24-byte records per tick,
    produce bit-identical velocities, patches and states to one context holding everything -- over three ticks with the
    state update applied on the device in between (tick -> compute_updates -> apply_updates -> gather -> tick)."""
    hz, cw = 20, 3
    p, cost, a, ms = cases.update_case(4242, hz)
    n = len(a["radius"])
    rng = np.random.default_rng(5)
    a["vdes"] = rng.normal(size=(n, 2)).astype(np.float32)
    a["vdes"] /= np.linalg.norm(a["vdes"], axis=1, keepdims=True)
    a["has_los"] = (rng.random(n) < 0.2).astype(np.uint32)
    rec, fl = capi.pack_agents(a)
    moving = (a["state"] != 2) & (a["state"] != 4)

    one = capi.Nav(0)
    _setup(one, cost, cw)
    one.agents_upload(rec, fl, hz)
    one.agents_upload_movestate(ms)
    navs = [capi.Nav(d) for d in _devices(world)]
    for nv in navs:
        _setup(nv, cost, cw)
    grp = capi.Group(navs)
    ranges = [capi.mgpu_shard_range(n, r, world) for r in range(world)]
    assert ranges[0][0] == 0 and ranges[-1][1] == n and all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
    for nv, (lo, hi) in zip(navs, ranges):
        nv.agents_upload_shard(rec[lo:hi], lo, hi, n, fl, hz)
    grp.gather()
    for nv, (lo, hi) in zip(navs, ranges):
        nv.agents_upload_movestate(ms[lo:hi])
    try:
        for tick in range(3):
            work = np.nonzero(moving)[0].astype(np.uint32)
            one.agents_set_work(work)
            one.agents_tick(0)
            v1 = one.agents_read_velocities(len(work))
            one.agents_compute_updates()
            p1 = one.agents_read_patches(len(work))
            one.agents_apply_updates()
            one.agents_rebuild_index()
            vs, ps = [], []
            for nv, (lo, hi) in zip(navs, ranges):
                w = work[(work >= lo) & (work < hi)]
                nv.agents_set_work(w)
                nv.agents_tick(0)
            for nv, (lo, hi) in zip(navs, ranges):
                w = work[(work >= lo) & (work < hi)]
                vs.append(nv.agents_read_velocities(len(w)))
                nv.agents_compute_updates()
                ps.append(nv.agents_read_patches(len(w)))
                nv.agents_apply_updates()
            grp.gather()
            vN, pN = np.concatenate(vs), np.concatenate(ps)
            assert (vN.view(np.uint32) == v1.view(np.uint32)).all(), "tick %d: velocities differ between 1 and %d contexts" % (tick, world)
            assert pN.tobytes() == p1.tobytes(), "tick %d: state patches differ" % tick
            a1, m1 = one.agents_read_state(n)
            aN = np.concatenate([nv.agents_read_state(hi - lo)[0] for nv, (lo, hi) in zip(navs, ranges)])
            assert aN.tobytes() == a1.tobytes(), "tick %d: entity state differs" % tick
            moving = (a1["state"] != 2) & (a1["state"] != 4)
        assert moving.sum() < (a["state"] != 2).sum(), "the scenario should see arrivals"
    finally:
        grp.close()
        for nv in navs:
            nv.close()
        one.close()


def test_shard_upload_same_flocks_flag(pf):
    """PFNAV_UPLOAD_SAME_FLOCKS re-upload (positions / velocities / states only) gives the same tick as a full upload"""
    cw = 1
    p, cost, a = cases.agent_case(cw, 900, 2, 77, 0.02, 2.4)
    rng = np.random.default_rng(8)
    a["vdes"] = rng.normal(size=(900, 2)).astype(np.float32)
    a["vdes"] /= np.linalg.norm(a["vdes"], axis=1, keepdims=True)
    a["has_los"] = np.zeros(900, np.uint32)
    rec, fl = capi.pack_agents(a)
    work = np.nonzero((a["state"] != 2) & (a["state"] != 4))[0].astype(np.uint32)
    nav = capi.Nav(0)
    try:
        nav.map_create(cw, cw, 1); nav.map_upload_layer(0, cost)
        nav.agents_upload_shard(rec, 0, 900, 900, fl, 20)
        rec2 = rec.copy()
        rec2["pos"] += rng.normal(scale=0.3, size=(900, 2)).astype(np.float32)
        rec2["prev_pos"] = rec2["pos"] - rec2["velocity"]
        nav.agents_upload_shard(rec2, 0, 900, 900, fl, 20, capi.UPLOAD_SAME_FLOCKS)
        nav.agents_set_work(work); nav.agents_tick(0)
        v_fast = nav.agents_read_velocities(len(work))
        nav.agents_upload(rec2, fl, 20)
        nav.agents_set_work(work); nav.agents_tick(0)
        v_full = nav.agents_read_velocities(len(work))
        assert (v_fast.view(np.uint32) == v_full.view(np.uint32)).all()
    finally:
        nav.close()


def test_pool_lru_eviction(pf, pforacle):
    """a full pool evicts the least recently used (dest, chunk) slots instead of failing (fieldcache.c:59-71), never a
    slot of the batch being written; evicted entries read as absent; a request that cannot fit fails without side effects"""
    cw = ch = 3
    p = cases.noise_map(cw, ch, 95, 0.05)
    cost = synth.cost_from_pathable(p, cw, ch)
    nav = capi.Nav(0)
    try:
        nav.map_create(cw, ch, 1); nav.map_upload_layer(0, cost); nav.map_build_nav(0)
        liid = nav.local_islands(0)
        om = pforacle.OracleMap(cw, ch, cost, None, liid)
        rng = np.random.default_rng(3)
        tiles = synth.random_passable_tiles(cost, 3, rng)
        targets = np.array([[int(t[0]) // cw, int(t[0]) % cw, int(t[1]), int(t[2])] for t in tiles], np.int32)
        nav.pool_create(3, 2 * cw * ch)                 # room for two destinations' field sets, three will be requested
        nchunks = []
        for d in range(2):
            nav.pool_request_goals(np.array([d], np.int32), targets[d:d + 1])
            nchunks.append(sum(nav.pool_get(d, (c // cw, c % cw))[0] is not None for c in range(cw * ch)))
        assert min(nchunks) >= 5
        # destination 0 is read by a tick (device-side LRU stamp), destination 1 is not
        a = synth.make_agents(cost, cw, ch, 64, 1, 11, radius=1.0, spacing=3.0)
        rec, fl = capi.pack_agents(a)
        fl["dest"] = 0
        nav.agents_upload(rec, fl, 20)
        nav.agents_set_work(np.arange(64, dtype=np.uint32))
        nav.agents_tick(capi.TICK_VDES_FROM_POOL)
        nav.agents_read_velocities(64)
        nav.pool_request_goals(np.array([2], np.int32), targets[2:3])       # must evict: the pool is nearly full
        have = [[nav.pool_get(d, (c // cw, c % cw))[0] is not None for c in range(cw * ch)] for d in range(3)]
        assert sum(have[2]) >= 5, "the new destination did not get its fields"
        assert sum(have[1]) < nchunks[1], "the never-read destination should have lost slots first"
        touched = np.unique(np.floor(a["pos"][:, 1] / 256).astype(int) * cw + np.floor(-a["pos"][:, 0] / 256).astype(int))
        assert all(have[0][c] for c in touched), "slots read by the last tick were evicted before untouched ones"
        # surviving and new fields still hold the right bytes
        fr, fc, fw, lr, lc = nav.plan_goal(tuple(int(v) for v in targets[2]))
        exp = {}
        for w in range(int(fw.max()) + 1):
            sel = np.nonzero(fw == w)[0]
            base = np.stack([exp.get(int(fc[i]), np.zeros((64, 64), np.uint8)) for i in sel])
            out = om.flow_fields_update(fr[sel], inout=base)
            for k, i in enumerate(sel):
                exp[int(fc[i])] = out[k]
        for c, f in exp.items():
            assert (nav.pool_get(2, (c // cw, c % cw))[0] == f).all()
        # a batch larger than the whole pool fails up front and leaves the pool untouched (nothing half-published)
        nav.pool_create(3, 4)
        with pytest.raises(capi.PfnavError):
            nav.pool_request_goals(np.arange(3, dtype=np.int32), targets)
        assert all(nav.pool_get(d, (c // cw, c % cw))[0] is None for d in range(3) for c in range(cw * ch))
    finally:
        nav.close()
