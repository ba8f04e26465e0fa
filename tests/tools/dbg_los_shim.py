"""debug: LOS of a blocked destination tile -- nav vs port vs reference vs shim"""
import sys, os, importlib.util, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases, pfref, pforacle
capi, synth = cases.capi, cases.synth
spec = importlib.util.spec_from_file_location("pfref_shim", os.path.join(ROOT, "oracle", "pfref.py"))
shim = importlib.util.module_from_spec(spec); spec.loader.exec_module(shim)
shim.LIB_PATH = os.path.join(ROOT, "oracle", "_ref", "libpfref_shim.so"); shim.lib()
cw = ch = 3
p = cases.noise_map(cw, ch, 8181, 0.08)
maps = [m.RefMap(cw, ch, p) for m in (pfref, shim)]
cost = maps[0].cost_base()
rng = np.random.default_rng(8181)
for _ in range(40):
    x, z, r = float(-rng.uniform(10, cw * 256 - 10)), float(rng.uniform(10, ch * 256 - 10)), float(rng.uniform(2, 9))
    for m in maps:
        m.blockers_incref(x, z, r)
for m in maps:
    m.update()
blk, liid = maps[0].blockers(), maps[0].local_islands()
td = (2, 0, 28, 30)
e = maps[0].los((2, 0), td)
s1 = maps[1].los((2, 0), td)
print("shim direct == ref:", (s1 == e).all(), (s1 != e).sum())
nav = capi.Nav(0)
nav.map_create(cw, ch, 1); nav.map_upload_layer(0, cost, blk, liid)
g = nav.los_fields_create(capi.los_req((2, 0), td))[0]
print("nav == ref:", (g == e).all(), (g != e).sum())
om = pforacle.OracleMap(cw, ch, cost, blk, liid)
print("port == ref:", (om.los_fields_create(capi.los_req((2, 0), td))[0] == e).all())
src, dst = (-502.0, 494.0), (-122.0, 626.0)
r0, r1 = maps[0].request_path(src, dst), maps[1].request_path(src, dst)
print("request", r0, r1)
for c in range(9):
    l0, l1 = maps[0].fc_los(r0[1], (c // 3, c % 3)), maps[1].fc_los(r1[1], (c // 3, c % 3))
    if l0 is not None:
        print(c, "los equal", (l0 == l1).all(), (l0 != l1).sum(), "ref cache == ref direct" if c != 6 else (l0 == e).all(), "shim cache == shim direct", (l1 == s1).all() if c == 6 else "")
s2 = maps[1].los((2, 0), td)
print("shim direct again == ref:", (s2 == e).all(), (s2 != e).sum())
