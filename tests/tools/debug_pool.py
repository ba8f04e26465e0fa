#!/usr/bin/env python3
"""GPU debug: batched pfnav_pool_request_goals vs the oracle port executing the same plan."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bench, pforacle
pf = importlib.import_module("permafrost-engine_b200"); capi, synth = pf.capi, pf.synth
W = bench.build_workload(pf, 1, 0)
cost = W["cost"]; a = W["agents"]
nav = capi.Nav(0); nav.map_create(16, 16, 1); nav.map_upload_layer(0, cost); nav.map_build_nav(0)
liid = nav.local_islands(0)
om = pforacle.OracleMap(16, 16, cost, None, liid)
goals = np.array([tuple(int(v) for v in a["flock_target_tile"][f]) for f in range(16)], np.int32)
nav.pool_create(16, 16 * 256)
mode = sys.argv[1] if len(sys.argv) > 1 else "batched"
if mode == "batched":
    print(nav.pool_request_goals(np.arange(16, dtype=np.int32), goals))
else:
    for d in range(16): nav.pool_request_goal(d, tuple(int(v) for v in goals[d]))
bad = 0
for d in range(16):
    fr, fc, fw, lr, lc = nav.plan_goal(tuple(int(v) for v in goals[d]))
    fields = {}
    for w in range(int(fw.max()) + 1):
        sel = np.nonzero(fw == w)[0]
        base = np.stack([fields.get(int(fc[i]), np.zeros((64, 64), np.uint8)) for i in sel])
        out = om.flow_fields_update(fr[sel], inout=base)
        for k, i in enumerate(sel): fields[int(fc[i])] = out[k]
    los = om.los_fields_create(lr)
    lmap = {int(lc[k]): los[k] for k in range(len(lr))}
    nb = nl = nm = 0
    for c in range(256):
        f, l, _ = nav.pool_get(d, (c // 16, c % 16))
        if (f is None) != (c not in fields): nm += 1; continue
        if f is not None and (f != fields[c]).any(): nb += 1
        if l is not None and c in lmap and (l != lmap[c]).any(): nl += 1
    print("dest", d, "flow-mismatch chunks", nb, "los-mismatch", nl, "presence-mismatch", nm, "waves", int(fw.max()) + 1)
    bad += nb + nl + nm
print("TOTAL BAD", bad)
