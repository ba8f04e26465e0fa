"""Device-side blocker reference counting (k_blockers_circles / k_blockers_tiles / k_chunks_finish) against the host
restatement of N_BlockersIncref / N_BlockersDecref / N_Update on a host-only context -- which tests/test_oracle.py pins
on the compiled reference (counts, faction counts, local islands, edge states). Bit-exact: counts of all 12 layers,
per-faction counts, local islands, and the planner's view after pfnav_map_commit (routes requested on both)."""
import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu
capi, synth = cases.capi, cases.synth


def _pair(cw, ch, nlayers, seed):
    cost = synth.cost_from_pathable(synth.make_map(cw, ch, seed), cw, ch)
    navs = [capi.Nav(0), capi.Nav(hostonly=True)]
    for nav in navs:
        nav.map_create(cw, ch, nlayers)
        for layer in range(nlayers):
            nav.map_upload_layer(layer, cost)
            nav.map_build_nav(layer)
    return cost, navs


def _same_state(dev, host, nlayers, check_factions=True):
    for layer in range(nlayers):
        assert (dev.blockers(layer) == host.blockers(layer)).all(), layer
        _, dblk, dliid = dev.map_get_layer(layer)
        assert (dblk == host.blockers(layer)).all(), layer              # the device grid itself, not only the mirror
        assert (dliid == host.local_islands(layer)).all(), layer
        assert (dev.local_islands(layer) == host.local_islands(layer)).all(), layer
        if check_factions:
            assert (dev.faction_counts(layer) == host.faction_counts(layer)).all(), layer


@pytest.mark.parametrize("seed", [3, 11])
def test_circle_ops_all_layers_vs_host(seed):
    cw, ch, nlayers = 3, 2, 12
    rng = np.random.default_rng(seed)
    p, (dev, host) = _pair(cw, ch, nlayers, 0x77 + seed)
    live = []
    for rnd in range(4):
        for _ in range(120):
            # anywhere on the map incl. the borders (clipped windows) and a few just outside (no tiles at all)
            x = float(-rng.uniform(-6, cw * 256 + 6)); z = float(rng.uniform(-6, ch * 256 + 6))
            r = float(rng.choice([0.0, 1.5, 3.0, 6.0, 9.5, 17.0, 47.9, 48.0, 60.0]))   # 60: wider than the bit window -> tile list
            f = int(rng.choice([-1, 0, 3, 14, 15]))
            flags = int(rng.choice([capi.FLAG_MOVABLE, capi.FLAG_MOVABLE | capi.FLAG_AIR, 0]))
            for nav in (dev, host):
                nav.blockers_incref(x, z, r, f, flags)
            live.append((x, z, r, f, flags))
        # take half of them away again (and a few twice in one batch: leave + re-enter inside one commit)
        rng.shuffle(live)
        gone, live = live[:len(live) // 2], live[len(live) // 2:]
        for (x, z, r, f, flags) in gone:
            for nav in (dev, host):
                nav.blockers_decref(x, z, r, f, flags)
        for (x, z, r, f, flags) in live[:10]:
            for nav in (dev, host):
                nav.blockers_decref(x, z, r, f, flags); nav.blockers_incref(x, z, r, f, flags)
        nd_dev, nd_host = dev.map_commit(), host.map_commit()
        assert 0 < nd_dev <= nd_host            # ours: passable set changed; the reference's: any 0 <-> n transition
        _same_state(dev, host, nlayers)
    for nav in (dev, host):
        nav.close()


def test_batch_and_obb_vs_host():
    cw, ch, nlayers = 2, 2, 4
    rng = np.random.default_rng(5)
    p, (dev, host) = _pair(cw, ch, nlayers, 0x99)
    ops = np.zeros(600, capi.BLOCKER_OP)
    ops["x"] = -rng.uniform(2, cw * 256 - 2, 600); ops["z"] = rng.uniform(2, ch * 256 - 2, 600)
    ops["range"] = rng.choice([1.0, 6.0, 12.0], 600); ops["faction_id"] = rng.integers(0, 4, 600)
    ops["flags"] = capi.FLAG_MOVABLE; ops["delta"] = 1
    for nav in (dev, host):
        nav.blockers_batch(ops)
        for k in range(12):                     # building footprints, rasterised on the host for both
            cx, cz = -float(40 + 37 * k), float(60 + 31 * k)
            ang = 0.3 * k
            ux, uz = np.cos(ang) * 14.0, np.sin(ang) * 14.0
            vx, vz = -np.sin(ang) * 9.0, np.cos(ang) * 9.0
            corners = [(cx - ux - vx, cz - uz - vz), (cx + ux - vx, cz + uz - vz), (cx + ux + vx, cz + uz + vz), (cx - ux + vx, cz - uz + vz)]
            nav.blockers_obb(np.array(corners, np.float32), True, k % 3, 0)
    assert dev.map_commit() <= host.map_commit()
    _same_state(dev, host, nlayers)
    ops["delta"] = -1
    for nav in (dev, host):
        nav.blockers_batch(ops[::2])
    dev.map_commit(); host.map_commit()
    _same_state(dev, host, nlayers)
    for nav in (dev, host):
        nav.close()


def test_routes_after_device_commit_equal_host():
    """the host planner reads the mirrors the device path refreshes: same edge states, same routes, same pool invalidation"""
    cw, ch = 4, 4
    rng = np.random.default_rng(9)
    p, (dev, host) = _pair(cw, ch, 1, 0xC5)
    for nav in (dev, host):
        nav.route_build(0)
    pos = np.stack([-rng.uniform(20, cw * 256 - 20, 400), rng.uniform(20, ch * 256 - 20, 400)], axis=1).astype(np.float32)
    for step in range(3):
        ops = np.zeros(len(pos), capi.BLOCKER_OP)
        ops["x"] = pos[:, 0]; ops["z"] = pos[:, 1]; ops["range"] = 6.0; ops["flags"] = capi.FLAG_MOVABLE; ops["delta"] = 1
        if step:
            old = ops.copy(); old["x"] = prev[:, 0]; old["z"] = prev[:, 1]; old["delta"] = -1
            ops = np.concatenate([old, ops])
        for nav in (dev, host):
            nav.blockers_batch(ops)
            nav.map_commit()
        assert (dev.local_islands(0) == host.local_islands(0)).all()
        assert (dev.route_islands(0) == host.route_islands(0)).all()
        for _ in range(40):
            src = (float(-rng.uniform(5, cw * 256 - 5)), float(rng.uniform(5, ch * 256 - 5)))
            dst = (float(-rng.uniform(5, cw * 256 - 5)), float(rng.uniform(5, ch * 256 - 5)))
            a, b = dev.route_request_path(src, dst), host.route_request_path(src, dst)
            assert type(a) is type(b)
            if isinstance(a, tuple):
                for u, v in zip(a, b):
                    assert np.array_equal(np.asarray(u), np.asarray(v)), (src, dst)
            else:
                assert a == b
        prev = pos.copy()
        pos = pos + rng.integers(-1, 2, size=pos.shape).astype(np.float32) * 4.0
        pos[:, 0] = np.clip(pos[:, 0], -(cw * 256 - 20), -20); pos[:, 1] = np.clip(pos[:, 1], 20, ch * 256 - 20)
    for nav in (dev, host):
        nav.close()
