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
