#!/usr/bin/env python3
"""Micro-benchmark (GPU box): flow-field and LOS-field kernels alone, device-resident requests,
CUDA-event timing, L2 flushed between iterations. Prints one JSON line per variant."""
import importlib, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
pf = importlib.import_module("permafrost-engine_b200")
capi, synth = pf.capi, pf.synth

def main():
    cw = ch = 16
    p = synth.make_map(cw, ch, 0x5EED0001)
    cost = synth.cost_from_pathable(p, cw, ch)
    nav = capi.Nav(0)
    nav.map_create(cw, ch, 1); nav.map_upload_layer(0, cost); nav.map_build_nav(0)
    rng = np.random.default_rng(0)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    tiles = synth.random_passable_tiles(cost, n, rng)
    tile_reqs = np.concatenate([capi.tile_req((int(t[0]) // cw, int(t[0]) % cw), (int(t[1]), int(t[2]))) for t in tiles])
    los_reqs = np.concatenate([capi.los_req((int(t[0]) // cw, int(t[0]) % cw), (int(t[0]) // cw, int(t[0]) % cw, int(t[1]), int(t[2]))) for t in tiles])
    # portal requests: the planner's own requests for a few goals, repeated
    pr = []
    for t in tiles[:8]:
        fr, fc, fw, lr, lc = nav.plan_goal((int(t[0]) // cw, int(t[0]) % cw, int(t[1]), int(t[2])))
        pr.append(fr[(fw == 0) & (fr["target_type"] == capi.TARGET_PORTAL)])
    portal_reqs = np.concatenate(pr)
    portal_reqs = np.concatenate([portal_reqs] * (n // len(portal_reqs) + 1))[:n]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    _stream = torch.cuda.Stream(); torch.cuda.set_stream(_stream)     # NULL would name the context's own stream
    st = _stream.cuda_stream
    out = torch.zeros((n, 4096), dtype=torch.uint8, device="cuda")
    def timeit(fn, iters=5):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        ms = []
        for _ in range(iters):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
        return float(np.median(ms))
    for name, reqs, bytes_per in (("flow_tile", tile_reqs, 16384), ("flow_portal", portal_reqs, 24704)):
        d = torch.from_numpy(reqs.view(np.uint8)).cuda()
        for tma in (0, 1):
            nav.set_tma(tma)
            ms = timeit(lambda: nav.flow_fields_update_dev(d.data_ptr(), n, out.data_ptr(), st))
            print(json.dumps({"kernel": "k_flow_unit", "case": name, "tma": tma, "n": n, "ms": ms, "fields_per_s": n / ms * 1e3,
                              "alg_GBps": n * bytes_per / ms / 1e6}), flush=True)
        ms = timeit(lambda: nav.flow_fields_update_dev(d.data_ptr(), n, out.data_ptr(), st, general=True), iters=3)
        print(json.dumps({"kernel": "k_flow_general", "case": name, "n": n, "ms": ms, "fields_per_s": n / ms * 1e3}), flush=True)
    nl = min(n, 4096)
    d = torch.from_numpy(los_reqs[:nl].view(np.uint8)).cuda()
    for v in (0,):
        ms = timeit(lambda: nav.los_fields_create_dev(d.data_ptr(), nl, out.data_ptr(), [0, nl], st), iters=3)
        print(json.dumps({"kernel": "k_los" + ("2" if v else ""), "case": "destination chunk", "n": nl, "ms": ms, "fields_per_s": nl / ms * 1e3,
                          "alg_GBps": nl * 16512 / ms / 1e6}), flush=True)
        ms1 = timeit(lambda: nav.los_fields_create_dev(d.data_ptr(), 1, out.data_ptr(), [0, 1], st), iters=3)
        print(json.dumps({"kernel": "k_los" + ("2" if v else ""), "case": "single field latency", "ms": ms1}), flush=True)

def goals():
    """the bench's field phase alone: 16 goals, dense plan, flow waves + LOS chain into the pool"""
    cw = ch = 16
    p = synth.make_map(cw, ch, 0x5EED0001)
    cost = synth.cost_from_pathable(p, cw, ch)
    nav = capi.Nav(0)
    nav.map_create(cw, ch, 1); nav.map_upload_layer(0, cost); nav.map_build_nav(0)
    rng = np.random.default_rng(3)
    tiles = synth.random_passable_tiles(cost, 16, rng)
    targets = np.array([[int(t[0]) // cw, int(t[0]) % cw, int(t[1]), int(t[2])] for t in tiles], np.int32)
    nav.pool_create(16, 16 * 256)
    _stream = torch.cuda.Stream(); torch.cuda.set_stream(_stream)
    st = _stream.cuda_stream
    for _ in range(3): nf, nl = nav.pool_request_goals(np.arange(16, dtype=np.int32), targets, 0, st)
    torch.cuda.synchronize()
    nav.profile_enable(True)
    ms = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); nav.pool_request_goals(np.arange(16, dtype=np.int32), targets, 0, st); nav.fields_join(st); b.record(); torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    prof = nav.profile_read()
    print(json.dumps({"case": "16 goals dense", "flow": nf, "los": nl, "ms_total_median": float(np.median(ms)),
                      "prof_flow_ms": prof["flow"][0] / 5, "prof_los_ms": prof["los"][0] / 5}), flush=True)


def region():
    """formation-sized batches of 96 x 96 cell arrival fields (formation.c:3152) on the C2 map, device-resident"""
    cw = ch = 16
    p = synth.make_map(cw, ch, 0x5EED0001)
    cost = synth.cost_from_pathable(p, cw, ch)
    nav = capi.Nav(0)
    nav.map_create(cw, ch, 1); nav.map_upload_layer(0, cost); nav.map_build_nav(0)
    rng = np.random.default_rng(5)
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 148 * 16
    dim = 96
    rec = np.zeros(n, capi.REGION_REQ)
    c = rng.integers(48, cw * 64 - 48, (n, 2))
    t = c + rng.integers(-40, 40, (n, 2))
    rec["center_r"], rec["center_c"] = c[:, 0], c[:, 1]
    rec["seed_off"] = np.arange(n); rec["seed_n"] = 1
    rec["flags"] = capi.REGION_CREATE | capi.REGION_CELL
    _stream = torch.cuda.Stream(); torch.cuda.set_stream(_stream)
    st = _stream.cuda_stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    d_req = torch.from_numpy(rec.view(np.uint8)).cuda()
    d_seed = torch.from_numpy(t.astype(np.int32)).cuda()
    d_ov = torch.zeros(2, dtype=torch.int32, device="cuda")
    out = torch.zeros((n, dim * dim // 2), dtype=torch.uint8, device="cuda")
    def run():
        nav.region_fields_dev(dim, d_req.data_ptr(), n, d_seed.data_ptr(), d_ov.data_ptr(), out.data_ptr(), st)
    for _ in range(3): run()
    torch.cuda.synchronize()
    ms = []
    for _ in range(5):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(); b.record(); torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    m = float(np.median(ms))
    print(json.dumps({"kernel": "k_region_fields", "case": "cell arrival 96x96, create", "n": n, "ms": m, "fields_per_s": n / m * 1e3,
                      "reached_frac": float((out != 0).float().mean().item())}), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "goals":
        goals(); sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "region":
        region(); sys.exit(0)
    main()
