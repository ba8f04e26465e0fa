"""Device-side AStar_PortalGraphPath (k_portal_graph_path) against the host planner's own routine -- the one
pfnav_route_request_path runs and tests/test_oracle.py pins on the compiled reference (ff_ids, field sets and LOS chains of
random requests): status, hop count, cost bits and every hop (chunk, portal, local island) must be identical, on maps with
rivers / fords, cut-off islands and committed dynamic obstacles (blocked edges)."""
import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu
capi, synth = cases.capi, cases.synth


def _searches(nav, cost, cw, ch, rng, n):
    """random (start tile, end tile, finish portal of the end chunk) triples; starts / ends on any tile, incl. impassable ones"""
    nports = [len(nav.portals_of(c)) for c in range(cw * ch)]
    req = np.zeros((n, 8), np.int32)
    k = 0
    while k < n:
        ec = int(rng.integers(0, cw * ch))
        if nports[ec] == 0:
            continue
        req[k] = [int(rng.integers(0, cw * ch)), int(rng.integers(0, 64)), int(rng.integers(0, 64)),
                  ec, int(rng.integers(0, 64)), int(rng.integers(0, 64)), ec, int(rng.integers(0, nports[ec]))]
        k += 1
    return req


@pytest.mark.parametrize("seed,cw,ch,blockers", [(1, 4, 4, 0), (2, 6, 3, 300), (3, 8, 8, 800)])
def test_portal_graph_paths_device_equals_host(seed, cw, ch, blockers):
    rng = np.random.default_rng(seed)
    cost = synth.cost_from_pathable(synth.make_map(cw, ch, 0xA57A + seed, frac_blocked=0.22), cw, ch)
    nav = capi.Nav(0)
    nav.map_create(cw, ch, 1); nav.map_upload_layer(0, cost); nav.map_build_nav(0); nav.route_build(0)
    ports = nav.portals(0)
    nav.portals_of = lambda c: ports[(ports[:, 0] * cw + ports[:, 1]) == c]
    if blockers:
        ops = np.zeros(blockers, capi.BLOCKER_OP)
        ops["x"] = -rng.uniform(4, cw * 256 - 4, blockers); ops["z"] = rng.uniform(4, ch * 256 - 4, blockers)
        ops["range"] = rng.choice([3.0, 6.0, 10.0], blockers); ops["flags"] = capi.FLAG_MOVABLE; ops["delta"] = 1
        nav.blockers_batch(ops)
        assert nav.map_commit() > 0
    req = _searches(nav, cost, cw, ch, rng, 600)
    st_d, cost_d, hops_d = nav.route_graph_paths(req, 256, True)
    st_h, cost_h, hops_h = nav.route_graph_paths(req, 256, False)
    assert (st_h >= 0).all() and (st_d == st_h).all(), np.nonzero(st_d != st_h)[0][:10]
    assert (cost_d.view(np.uint32) == cost_h.view(np.uint32)).all()
    for i in range(len(req)):
        assert np.array_equal(hops_d[i], hops_h[i]), (i, req[i])
    found = int((st_h == 1).sum())
    assert found >= 60, found                      # the comparison is not vacuous
    assert max(len(h) for h in hops_h) >= max(cw, ch)
    nav.close()
