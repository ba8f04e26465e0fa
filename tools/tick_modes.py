#!/usr/bin/env python3
"""time pfnav_agents_tick on the C2 population in the three velocity-update modes (no fields in flight)"""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
pf = importlib.import_module("permafrost-engine_b200")
capi = pf.capi
W = bench.build_workload(pf, 1, 0)
nav = capi.Nav(0)
nav.map_create(16, 16, 1); nav.map_upload_layer(0, W["cost"]); nav.map_build_nav(0)
a = W["agents"]
tgt = a["flock_target"][a["flock_of"]]; d = tgt - a["pos"]; d /= np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-3)
rec = W["rec"].copy(); rec["vdes"] = d.astype(np.float32)
nav.agents_upload(rec, W["flocks"], 20)
nav.agents_set_work(np.arange(len(rec), dtype=np.uint32))
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
for mode in (0, 2, 0, 2):
    nav.set_two_phase(mode)
    for _ in range(2): nav.agents_tick(0, st.cuda_stream)
    torch.cuda.synchronize()
    ms = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); nav.agents_tick(0, st.cuda_stream); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    print("mode", mode, "tick ms", np.round(ms, 3))
