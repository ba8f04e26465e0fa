"""offline fuzz: the repair chain of N_DesiredPointSeekVelocity (N_FlowFieldUpdateToNearestPathable / ...IslandToNearest),
port + host seed code vs the compiled reference, on more maps with blockers"""
import os, sys, time, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import cases, pfref, pforacle
import test_oracle as T
t0 = time.time(); tot = bad = 0
for seed in range(4000, 4008):
    cw, ch = [(2, 2), (3, 2), (2, 3), (3, 3)][seed % 4]
    rng = np.random.default_rng(seed)
    p = cases.synth.make_map(cw, ch, seed, frac_blocked=0.12, rivers=True); p[rng.random(p.shape) < [0.1, 0.25, 0.4][seed % 3]] = 0
    ref = pfref.RefMap(cw, ch, p)
    for _ in range(60):
        ref.blockers_incref(float(-rng.uniform(10, cw * 256 - 10)), float(rng.uniform(10, ch * 256 - 10)), float(rng.uniform(2, 16)), 0, 0)
    ref.update()
    Tq, K, A, B, E = cases.repair_case(ref, cw, ch, seed, per_chunk=3)
    om = pforacle.OracleMap(cw, ch, ref.cost_base(), ref.blockers(), ref.local_islands())
    got = T._port_repair(pforacle, om, ref.islands(), Tq, K, A, B)
    b = int((got != E).reshape(len(E), -1).any(axis=1).sum())
    tot += len(E); bad += b
    print("seed", seed, (cw, ch), "repairs", len(E), "bad", b, "%.0fs" % (time.time() - t0), flush=True)
    ref.close()
print("TOTAL", tot, "bad", bad)
