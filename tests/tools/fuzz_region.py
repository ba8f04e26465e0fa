"""offline fuzz (not part of the suite): port / host code / kernel-source emulation vs the compiled reference on many seeds"""
import os, sys, time, ctypes as C, numpy as np, subprocess, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import cases, pfref, pforacle
capi, synth = cases.capi, cases.synth
t0 = time.time()
so = "/tmp/libregion_emu_fuzz.so"
subprocess.run(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-pthread", os.path.join(ROOT, "tests", "emu", "region_emu.cpp"), "-o", so], check=True)
E = C.CDLL(so)
E.emu_region_fields.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
def P(a): return a.ctypes.data_as(C.c_void_p)
def image(a, cw, ch): return np.ascontiguousarray(a.reshape(ch, cw, 64, 64).transpose(0, 2, 1, 3).reshape(ch * 64, cw * 64))
def fmask(fac, cw, ch):
    m = np.zeros((cw * ch, 64, 64), np.uint16)
    for f in range(15): m |= (fac[:, f] > 0).astype(np.uint16) << f
    return image(m, cw, ch)
tot = bad = 0
for seed in range(1000, 1012):
    cw, ch = [(3, 3), (2, 4), (5, 2), (4, 4)][seed % 4]
    dim = [96, 32, 64, 128, 96, 48][seed % 6]
    if dim // 2 >= min(cw, ch) * 32:      # region_case's "interior" draws need room
        dim = 32
    rngm = np.random.default_rng(seed)
    p = synth.make_map(cw, ch, seed, frac_blocked=0.12, rivers=True); p[rngm.random(p.shape) < [0.05, 0.15, 0.3][seed % 3]] = 0
    _, blockers, wars, reqs = cases.region_case(seed, cw, ch, 36, dim) if cw == ch else (None, None, None, None)
    if reqs is None:
        # non-square: build the request list by hand
        rng = np.random.default_rng(seed); H, W = ch * 64, cw * 64
        blockers = [(float(-rng.uniform(10, cw * 256 - 10)), float(rng.uniform(10, ch * 256 - 10)), float(rng.uniform(2, 12)), int(rng.integers(0, 4))) for _ in range(60)]
        wars = [(0, 1), (0, 2), (3, 1)]; reqs = []
        for i in range(36):
            c_ = (int(rng.integers(0, H)), int(rng.integers(0, W))); half = dim // 2
            t_ = (int(rng.integers(max(c_[0] - half, 0), min(c_[0] + half - 1, H - 1) + 1)), int(rng.integers(max(c_[1] - half, 0), min(c_[1] + half - 1, W - 1) + 1)))
            reqs.append(dict(center=c_, target=t_, enemies=[0, 2, 6, 15][i % 4], overlay=None, start=None, want_fixup=(i % 2 == 0)))
    else:
        p = cases.noise_map(cw, ch, seed, 0.10)
    ref = pfref.RefMap(cw, ch, p)
    for a, b in wars: ref.set_war(a, b)
    for b in blockers: ref.blockers_incref(b[0], b[1], b[2], b[3], 0)
    ref.update()
    cost, blk, fac = ref.cost_base(), ref.blockers(), ref.factions()
    cases.region_pick_starts(reqs, cost, blk, cw, ch, seed, dim)
    om = pforacle.OracleMap(cw, ch, cost, blk, None, factions=fac)
    exp = [ref.cell_arrival_field(dim, q["target"], q["center"], q["enemies"], q["overlay"], q["start"]) for q in reqs]
    rec, sd, ov = capi.pack_region_reqs(reqs)
    sd2 = np.ascontiguousarray(np.concatenate([sd, np.zeros((1, 2), np.int32)])); ov2 = np.ascontiguousarray(np.concatenate([ov, np.zeros((1, 2), np.int32)]))
    got = np.zeros((len(reqs), dim, dim // 2), np.uint8)
    ci, bi, fi = image(cost, cw, ch), image(blk, cw, ch), fmask(fac, cw, ch)
    E.emu_region_fields(P(ci), P(bi), P(fi), cw * 64, ch * 64, dim, P(rec), len(rec), P(sd2), P(ov2), P(got), 0)
    b1 = b2 = 0
    for i, q in enumerate(reqs):
        e = om.region_field_create(dim, q["enemies"], 1, [q["target"]], q["center"], q["overlay"])
        if q["start"] is not None: e = om.region_field_fixup(dim, q["start"], q["center"], e, q["overlay"])
        b1 += int((e != exp[i]).any()); b2 += int((got[i] != exp[i]).any())
    # zone fields: every chunk, a few centres; host seeds + emulated kernel window
    zb = zt = 0
    if min(cw, ch) > 1:
        nav = capi.Nav(hostonly=True); nav.map_create(cw, ch, 1); nav.map_upload_layer(0, cost, blk)
        rng = np.random.default_rng(seed + 7)
        for k in range(4):
            centre = (int(rng.integers(0, ch * 64)), int(rng.integers(0, cw * 64))); rad = int(rng.choice([0, 1, 4, 9, 20, 60, 130]))
            recz = np.zeros(cw * ch, capi.REGION_REQ); sds = []
            for c in range(cw * ch):
                s_ = nav.zone_seeds((c // cw, c % cw), centre, rad)
                recz["center_r"][c], recz["center_c"][c] = c // cw, c % cw
                recz["seed_off"][c], recz["seed_n"][c] = sum(len(x) for x in sds), len(s_); recz["flags"][c] = capi.REGION_CREATE
                sds.append(s_)
            sdz = np.ascontiguousarray(np.concatenate(sds + [np.zeros((1, 2), np.int32)]).astype(np.int32))
            gz = np.full((cw * ch, 64, 64), 0xEE, np.uint8)
            E.emu_region_fields(P(ci), P(bi), P(np.zeros_like(bi)), cw * 64, ch * 64, 128, P(recz), len(recz), P(sdz), P(ov2), P(gz), 1)
            for c in range(cw * ch):
                ez = ref.flow_field_zone((c // cw, c % cw), centre, rad)
                zt += 1; zb += int((gz[c] != ez).any())
        nav.close()
    tot += len(reqs) + zt; bad += b1 + b2 + zb
    print("seed", seed, (cw, ch), "dim", dim, "region: port bad", b1, "kernel-source bad", b2, "of", len(reqs), "| zone bad", zb, "of", zt, "| %.0fs" % (time.time() - t0), flush=True)
    ref.close()
print("TOTAL", tot, "bad", bad)
