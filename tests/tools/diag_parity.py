"""Diagnostic (not part of the suite): error distribution of the CUDA velocity pass against the compiled
reference on the golden populations and on seeded samples of the bench's own C2 / C3 populations.
Writes gpurun_out/diag_parity.json.  TEST INFRASTRUCTURE (uses oracle/)."""
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
import pfref  # noqa: E402

pf = importlib.import_module("permafrost-engine_b200")
capi, synth = pf.capi, pf.synth
GOLD = os.path.join(ROOT, "tests", "golden")
out = {}


def stats(name, got, exp):
    e = cases.relerr(got, exp)
    bad = np.nonzero(e > 1e-4)[0]
    out[name] = dict(n=len(e), max=float(e.max()), n_bad=int(len(bad)), bad=[int(b) for b in bad[:64]],
                     bad_err=[float(e[b]) for b in bad[:64]], p999=float(np.quantile(e, 0.999)),
                     n_gt_1e6=int((e > 1e-6).sum()))
    print(name, out[name]["max"], out[name]["n_bad"], out[name]["n_gt_1e6"], flush=True)


def golden(nav):
    for name, cw in (("agents_1x1", 1), ("agents_dense", 1), ("agents_3x3", 3), ("update_hz20", 3), ("update_hz10", 3)):
        g = np.load(os.path.join(GOLD, name + ".npz"))
        a = {k[2:]: g[k] for k in g.files if k.startswith("a_")}
        a["vdes"] = np.zeros((len(a["radius"]), 2), np.float32); a["vdes"][g["work"]] = g["vdes"]
        a["has_los"] = np.zeros(len(a["radius"]), np.uint32); a["has_los"][g["work"]] = g["los"]
        rec, fl = capi.pack_agents(a)
        nav.map_create(cw, cw, 1); nav.map_upload_layer(0, g["cost"])
        hz = int(g["hz"]) if "hz" in g.files else 20
        nav.agents_upload(rec, fl, hz)
        nav.agents_set_work(g["work"])
        for mode in (0, 2):
            nav.set_two_phase(mode)
            nav.agents_tick(0)
            vel = nav.agents_read_velocities(len(g["work"]))
            vpref, _, _ = nav.agents_read_debug(len(g["work"]))
            stats("%s/mode%d/vel" % (name, mode), vel, g["vel"])
            if "vpref" in g.files:
                stats("%s/mode%d/vpref" % (name, mode), vpref, g["vpref"])
        nav.set_two_phase(1)


def scale(nav, workload, nsample=10000):
    import bench
    bench.set_workload(workload)
    t0 = time.time()
    W = bench.build_workload(pf, 1, 0)
    a = W["agents"]
    n = W["n_total"]
    print(workload, "population built", time.time() - t0, flush=True)
    t0 = time.time()
    ref = pfref.RefMap(bench.CHUNKS, bench.CHUNKS, W["pathable"])
    out[workload + "/refmap_s"] = time.time() - t0
    print("refmap", time.time() - t0, flush=True)
    assert (ref.cost_base() == W["cost"]).all()
    rng = np.random.default_rng(7)
    work = np.sort(rng.choice(n, nsample, replace=False)).astype(np.uint32)
    d = a["flock_target"][a["flock_of"]] - a["pos"]
    vdes_all = (d / np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-6)).astype(np.float32)
    los_all = (rng.random(n) < 0.2).astype(np.uint8)
    dest = np.arange(W["nflocks"], dtype=np.uint32)
    t0 = time.time()
    ref.agents_set(a["pos"], a["prev_pos"], a["vel"], a["radius"], a["max_speed"], a["state"], a["flags"],
                   a["flock_of"], a["flock_target"], dest, hz=20)
    ref.work_set(work, vdes_all[work], los_all[work], a["speed"][work])
    print("agents_set", time.time() - t0, flush=True)
    evel, secs = ref.velocity_work(os.cpu_count())
    evpref = ref.vpref()
    out[workload + "/ref_secs"] = secs
    print("ref velocity_work", secs, flush=True)
    nav.map_create(bench.CHUNKS, bench.CHUNKS, 1); nav.map_upload_layer(0, W["cost"])
    aa = dict(a); aa["vdes"] = vdes_all; aa["has_los"] = los_all.astype(np.uint32)
    rec, fl = capi.pack_agents(aa)
    nav.agents_upload(rec, fl, 20)
    nav.agents_set_work(work)
    for mode in (0, 2):
        nav.set_two_phase(mode)
        nav.agents_tick(0)
        vel = nav.agents_read_velocities(nsample)
        vpref, _, _ = nav.agents_read_debug(nsample)
        stats("%s/mode%d/vel" % (workload, mode), vel, evel)
        stats("%s/mode%d/vpref" % (workload, mode), vpref, evpref)
    nav.set_two_phase(1)
    # neighbour order on 300 sampled agents, both radii
    mism = 0
    for i in work[:300]:
        x, z = float(a["pos"][i, 0]), float(a["pos"][i, 1])
        for r, cap in ((10.0, 512), (30.0, 128)):
            g_ = nav.ents_in_circle(x, z, r, cap); e_ = ref.ents_in_circle(x, z, r, cap)
            if len(g_) != len(e_) or (g_ != e_).any():
                mism += 1
    out[workload + "/order_mismatches"] = mism
    print("order mismatches", mism, flush=True)
    nz = np.linalg.norm(evel, axis=1) == 0
    out[workload + "/zero_vel_frac"] = float(nz.mean())
    ref.close()


if __name__ == "__main__":
    nav = capi.Nav(0)
    golden(nav)
    for wl in sys.argv[1:] or ["C2", "C3"]:
        scale(nav, wl)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "diag_parity.json"), "w") as f:
        json.dump(out, f, indent=1)
    nav.close()
