"""Link-swap proof of the drop-in boundary (seam B2). oracle/_ref/libpfref_shim.so is the reference's own
nav.c + a_star.c + fieldcache.c + movement.c ... with src/navigation/field.c REPLACED by shim/field_pfnav.c, which exports
field.h's exact signatures on top of libpfnav.so (built by `make -C oracle shimref`, travels to the GPU box like
libpfref.so). The reference's n_request_path, field cache, N_DesiredPointSeekVelocity (with its on-miss chain and both
field repairs) and N_HasDestLOS then run UNCHANGED on GPU-built fields -- and must return exactly what the all-reference
library returns."""
import importlib.util
import os

import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu
capi, synth = cases.capi, cases.synth
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def both(pfref):
    path = os.path.join(ROOT, "oracle", "_ref", "libpfref_shim.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libpfref_shim.so not built (make -C oracle shimref)")
    spec = importlib.util.spec_from_file_location("pfref_shim", os.path.join(ROOT, "oracle", "pfref.py"))
    shim = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(shim)
    shim.LIB_PATH = path
    shim.lib()
    return pfref, shim


def test_reference_nav_on_gpu_fields(both):
    ref_mod, shim_mod = both
    cw = ch = 3
    p = cases.noise_map(cw, ch, 8181, 0.08)
    maps = [m.RefMap(cw, ch, p) for m in (ref_mod, shim_mod)]
    try:
        cost = maps[0].cost_base()
        rng = np.random.default_rng(8181)
        # dynamic obstacles first: blocked targets, cut-off islands -> the repair chain has work to do
        for _ in range(40):
            x, z, r = float(-rng.uniform(10, cw * 256 - 10)), float(rng.uniform(10, ch * 256 - 10)), float(rng.uniform(2, 9))
            for m in maps:
                m.blockers_incref(x, z, r)
        for m in maps:
            m.update()
        assert (maps[0].blockers() == maps[1].blockers()).all() and (maps[0].local_islands() == maps[1].local_islands()).all()
        # direct field calls through field.h
        tiles = np.argwhere(cost[4] != 255)
        t = tuple(int(v) for v in tiles[len(tiles) // 3])
        assert (maps[0].flow_tile((1, 1), t) == maps[1].flow_tile((1, 1), t)).all()
        lr = cases.los_case(cost, cw, ch, 5, ntargets=3)
        assert (cases.ref_los_batch(maps[0], lr) == cases.ref_los_batch(maps[1], lr)).all()
        # n_request_path: the reference's planner, cache and merge logic on top of shim-built fields
        pairs = cases.route_pairs(cost, cw, ch, 4, 24)
        nok = 0
        for src, dst in pairs:
            res = [m.request_path(src, dst) for m in maps]
            assert res[0] == res[1], (src, dst, res)
            ok, did = res[0]
            nok += ok
            if not ok:
                continue
            for c in range(cw * ch):
                f0, id0 = maps[0].fc_flow(did, (c // cw, c % cw)); f1, id1 = maps[1].fc_flow(did, (c // cw, c % cw))
                assert (f0 is None) == (f1 is None) and id0 == id1, (src, dst, c)
                assert f0 is None or (f0 == f1).all(), (src, dst, c)
                l0, l1 = maps[0].fc_los(did, (c // cw, c % cw)), maps[1].fc_los(did, (c // cw, c % cw))
                assert (l0 is None) == (l1 is None) and (l0 is None or (l0 == l1).all()), (src, dst, c)
        assert nok >= 12
        # N_DesiredPointSeekVelocity / N_HasDestLOS for entities anywhere on the map, incl. blocked tiles and walls:
        # on-miss requests + N_FlowFieldUpdateToNearestPathable / N_FlowFieldUpdateIslandToNearest through the shim
        for src, dst in pairs[:6]:
            ok, did = maps[0].request_path(src, dst)
            if not ok:
                continue
            pos = np.stack([-rng.uniform(2, cw * 256 - 2, 400), rng.uniform(2, ch * 256 - 2, 400)], 1).astype(np.float32)
            out = [m.desired_velocity(did, pos, pos, dst) for m in maps]
            assert (out[0][0] == out[1][0]).all(), np.nonzero((out[0][0] != out[1][0]).any(axis=1))[0][:10]
            assert (out[0][1] == out[1][1]).all()
        # arrival fields (nav.h:700-730) and a TARGET_ZONE chunk field
        img = synth.blocked_to_image(cost, cw, ch)
        pas = np.argwhere(img != 255)
        for k in range(6):
            ctr = tuple(int(v) for v in pas[rng.integers(len(pas))])
            tgt = (min(max(ctr[0] + int(rng.integers(-30, 30)), 0), ch * 64 - 1), min(max(ctr[1] + int(rng.integers(-30, 30)), 0), cw * 64 - 1))
            if abs(tgt[0] - ctr[0]) >= 48 or abs(tgt[1] - ctr[1]) >= 48:
                continue
            a0, a1 = [m.cell_arrival_field(96, tgt, ctr) for m in maps]
            assert (a0 == a1).all(), (k, ctr, tgt)
        ctr = tuple(int(v) for v in pas[len(pas) // 2])
        z0, z1 = [m.flow_field_zone((ctr[0] // 64, ctr[1] // 64), ctr, 9) for m in maps]
        assert (z0 == z1).all()
    finally:
        for m in maps:
            m.close()


# ------------------------------------------------------------------------------------------
# Seam B1: the reference's own GPU back-end hooks (render.h:619-691), compiled on libpfnav.so
# ------------------------------------------------------------------------------------------
B1_ENT = np.dtype([("dest", "<f4", 2), ("vdes", "<f4", 2), ("cell_pos", "<f4", 2), ("coh", "<f4", 2), ("align", "<f4", 2),
                   ("drag", "<f4", 2), ("pos", "<f4", 2), ("velocity", "<f4", 2), ("movestate", "<u4"), ("flock_id", "<u4"),
                   ("flags", "<u4"), ("speed", "<f4"), ("max_speed", "<f4"), ("radius", "<f4"), ("layer", "<u4"),
                   ("has_dest_los", "<u4"), ("ready", "<u4"), ("pad0", "<u4")])             # struct gpu_ent_desc, movement.c:350
B1_FLOCK = np.dtype([("ents", "<u4", 1024), ("nmembers", "<u4"), ("target", "<f4", 2)])      # struct gpu_flock_desc, :341
B1_RES = np.dtype([("chunk_w", "<i4"), ("chunk_h", "<i4"), ("tile_w", "<i4"), ("tile_h", "<i4"), ("field_w", "<f4"),
                   ("field_h", "<f4")])                                                     # struct map_resolution, tile.h:127


def test_b1_movement_backend_equals_direct_tick(pf):
    """shim/gl_movement_pfnav.c: R_GL_MoveUploadData / UpdateUniforms / DispatchWork / ReadNewVelocities called as
    move_submit_gpu_velocity_work calls them (movement.c:3943), with buffers laid out as move_upload_input packs them
    (:3790-3915), against pfnav_agents_tick on the same population (prev_pos = pos: the seam carries no prev_pos)."""
    import ctypes as C
    path = os.path.join(ROOT, "oracle", "_ref", "libpfnav_b1.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libpfnav_b1.so not built (make -C oracle shimb1)")
    L = C.CDLL(path)
    cw = 3
    p, cost, a = cases.agent_case(cw, 2000, 3, 515, 0.04, 2.4)
    n = len(a["radius"])
    a["prev_pos"] = a["pos"].copy()
    d = a["flock_target"][a["flock_of"]] - a["pos"]
    a["vdes"] = (d / np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-6)).astype(np.float32)
    rng = np.random.default_rng(5)
    a["has_los"] = (rng.random(n) < 0.3).astype(np.uint32)
    work = np.nonzero((a["state"] != 2) & (a["state"] != 4))[0].astype(np.uint32)
    # blockers under the arrived entities, as the engine would hold them
    blk = np.zeros((cw * cw, 64, 64), np.uint16)
    # ---- direct ----
    nav = capi.Nav(0)
    nav.map_create(cw, cw, 1); nav.map_upload_layer(0, cost, blk)
    rec, fl = capi.pack_agents(a)
    fl["dest"] = -1
    for hz in (20, 10):
        nav.agents_upload(rec, fl, hz); nav.agents_set_work(work); nav.agents_tick(0)
        want = nav.agents_read_velocities(len(work))
        # ---- through the seam ----
        ents = np.zeros(n, B1_ENT)
        ents["dest"] = a["flock_target"][a["flock_of"]]; ents["vdes"] = a["vdes"]; ents["pos"] = a["pos"]; ents["velocity"] = a["vel"]
        ents["movestate"] = a["state"]; ents["flock_id"] = a["flock_of"] + 1; ents["flags"] = a["flags"]
        ents["speed"] = a["speed"]; ents["max_speed"] = a["max_speed"]; ents["radius"] = a["radius"]; ents["has_dest_los"] = a["has_los"]
        flocks = np.zeros(len(a["flock_target"]), B1_FLOCK)
        flocks["target"] = a["flock_target"]
        for f in range(len(flocks)):          # member lists cut at 1024 like the engine's buffer (the shim ignores them)
            m = np.nonzero(a["flock_of"] == f)[0][:1024]
            flocks["ents"][f, :len(m)] = m + 1; flocks["nmembers"][f] = len(m)
        gpuids = (work + 1).astype(np.uint32)
        res = np.zeros(1, B1_RES); res["chunk_w"] = cw; res["chunk_h"] = cw; res["tile_w"] = 32; res["tile_h"] = 32
        map_pos = np.zeros(2, np.float32)
        sz = lambda v: C.byref(C.c_size_t(int(v)))
        ptr = lambda arr: arr.ctypes.data_as(C.c_void_p)
        costb = np.ascontiguousarray(cost, np.uint8)
        L.R_GL_MoveUploadData(ptr(gpuids), sz(len(gpuids)), ptr(ents), sz(ents.nbytes), ptr(flocks), sz(flocks.nbytes),
                              ptr(costb), sz(costb.nbytes), ptr(blk), sz(blk.nbytes))
        L.R_GL_MoveUpdateUniforms(ptr(res), ptr(map_pos), C.byref(C.c_int(hz)), C.byref(C.c_int(len(work))))
        L.R_GL_MoveDispatchWork(sz(n))
        done = C.c_int(0)
        L.R_GL_MovePollCompletion(C.byref(done))
        assert done.value == 1
        got = np.zeros((len(work), 2), np.float32)
        L.R_GL_MoveReadNewVelocities(ptr(got), sz(len(work)), sz(len(work)))
        L.R_GL_MoveInvalidateData()
        assert (got == want).all(), (hz, np.abs(got - want).max())
        assert np.abs(want).max() > 0.1
    L.R_GL_MoveClearState()
    nav.close()
