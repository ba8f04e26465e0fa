#!/usr/bin/env python3
"""Where the LOS phase of the C2 step goes: per-field trace of one pfnav_pool_request_goals call (16 goals x 256 chunks).
Prints the makespan, the critical dependency chain (field by field: wait, run, pops) and aggregate rates."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
pf = importlib.import_module("permafrost-engine_b200")
capi = pf.capi
W = bench.build_workload(pf, 1, 0)
nav = capi.Nav(0)
C = bench.CHUNKS
nav.map_create(C, C, 1); nav.map_upload_layer(0, W["cost"]); nav.map_build_nav(0)
goals = np.array([tuple(int(v) for v in W["agents"]["flock_target_tile"][f]) for f in range(16)], np.int32)
nav.pool_create(16, 16 * C * C)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
dests = np.arange(16, dtype=np.int32)
for _ in range(3):
    nav.pool_request_goals(dests, goals, 0, st.cuda_stream); nav.fields_join(st.cuda_stream)
torch.cuda.synchronize()
nav.los_trace(True)
nav.pool_request_goals(dests, goals, 0, st.cuda_stream); nav.fields_join(st.cuda_stream)
torch.cuda.synchronize()
tr = nav.los_trace(False, read_cap=16 * C * C).astype(np.int64)
prev = (tr[:, 3] >> 32) - 1
tr[:, 3] &= 0xFFFFFFFF
t0 = tr[:, 0].min()
take, ready, done, pops = tr[:, 0] - t0, tr[:, 1] - t0, tr[:, 2] - t0, tr[:, 3]
print("fields %d, non-empty %d, makespan %.3f ms" % (len(tr), (pops > 0).sum(), done.max() / 1e6))
run = done - ready
ne = pops > 0
print("non-empty fields: pops total %d, run time total %.2f ms, mean %.1f us/field, %.3f us/pop" %
      ((pops[ne] - 1).sum(), run[ne].sum() / 1e6, run[ne].mean() / 1e3, run[ne].sum() / 1e3 / max((pops[ne] - 1).sum(), 1)))
print("empty fields: %d, mean %.2f us each" % ((~ne).sum(), run[~ne].mean() / 1e3))
# critical chain: walk the real dependencies back from the last finisher
i = int(np.argmax(done)); chain = []
while i >= 0:
    chain.append(i)
    i = int(prev[i])
chain = chain[::-1]
print("critical chain (%d fields):" % len(chain))
tot_run = 0
for k in chain:
    tot_run += run[k]
    if pops[k] > 0 or run[k] > 5000:
        print("  field %5d  take %8.1f us  ready %8.1f us  done %8.1f us  run %7.1f us  pops %5d" %
              (k, take[k] / 1e3, ready[k] / 1e3, done[k] / 1e3, run[k] / 1e3, max(pops[k] - 1, 0)))
print("chain run time %.2f ms of makespan %.2f ms; empty fields on the chain: %d" % (tot_run / 1e6, done.max() / 1e6, sum(1 for k in chain if pops[k] == 0)))
# slack: for non-empty fields, time between dependency satisfied... (always 0 by construction) ; time between take and start when no dep
late = [(k, take[k]) for k in range(len(tr)) if pops[k] > 0]
print("latest take of a non-empty field: %.1f us" % (max(t for _, t in late) / 1e3))
# per-goal roots
roots = np.nonzero(prev < 0)[0]
print("roots:", [(int(r), int(max(pops[r]-1,0)), round(run[r]/1e3,1)) for r in roots])
