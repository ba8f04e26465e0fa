#!/usr/bin/env python3
"""bench.py -- the driver's measurement contract for the navigation / crowd-movement hot path.

    python bench.py --gpus N --steps K --warmup W [--workload C1|C2|C3|C4|C5]   (N > 1 via torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W

A "step" is one movement tick over one synthetic batch (all per GPU, weak scaling):
   C2 (default at N = 1, BASELINE.json configs[1]): 1024x1024-tile map (16x16 chunks), 100 k agents, 16 goals
   C1 64x64 tiles / 256 agents / 1 goal;  C3 1024^2 / 1 M agents / 64 goals
   C4 (default at N > 1, configs[3] at N = 8): 2048x2048-tile map, 500 k agents + 8 goals per GPU, every flock
      confined to its own cell of an 8x8 grid over the map, so the local density does not change with N
   C5 (configs[4] at N = 8): 1024^2 map, 62.5 k agents + 2 goals per GPU and 1 000 dynamic obstacles (replicated on every
      rank) that all move one tile per tick: N_BlockersDecref + N_BlockersIncref, N_Update, the invalidated fields of
      the rank's goals rebuilt, then the tick
One step =
   1. fields: the goals' field sets (flow waves + chained LOS, every chunk connected to the goal) built into the
      device field pool. `value` is measured COLD: the pool is cleared and every goal is re-planned on the host
      (pfnav_pool_clear + pfnav_pool_request_goals) INSIDE the timed region. The same step with the request plan
      resident on the device (`value_resident_plan`, the round-1 headline) and WARM with the fields cached, which is
      what a steady-state tick of the reference pays (`value_warm`), are timed right after and reported beside it.
   2. (N > 1) pfnav_mgpu_gather: ONE all-gather of the 24-byte neighbour records inside libpfnav.so (NCCL), then the
      position index over the whole population on every rank.
   3. the agent tick: desired velocity + LOS out of the pool, cohesion, boids steering, neighbours, ClearPath.
`value` = agent updates per second, everything resident in HBM; `e2e` = the same cold step through the host-buffer C ABI
(this rank's 64-byte entity records H2D from pinned memory, its velocities D2H, every step). `value_moving` is a
device-resident multi-tick run (tick -> entity_compute_update -> apply -> gather, positions advance, fields warm).

--impl reference times the reference's own CPU implementation (oracle/_ref, the unmodified sources compiled by
oracle/Makefile; the C port when that is absent) on the host cores, on a bounded sample of the same workload.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HZ = 20
ALG_BYTES_PER_AGENT = 360           # SURVEY.md 8d "Canonical figure"
ALG_BYTES_PER_FLOW_FIELD = 24_704   # TARGET_PORTAL: cost 4096 + blockers 8192 + islands 8192+128 + dirs 4096
ALG_BYTES_PER_LOS_FIELD = 16_512
ISSUE_PEAK_PER_SM_CLK = 4           # warp instructions per clock per SM (4 schedulers)

# per-GPU sizes; seed = SURVEY.md 8d's 0x5EED0000 + config#
WORKLOADS = {
    "C1": dict(chunks=1, agents=256, goals=1, radius=[3.0], spacing=4.0, seed=0x5EED0000, agent_seed=0x5EED0000, grid=None,
               desc="64x64-tile map (1 chunk), 256 agents in 1 flock, 1 goal"),
    "C2": dict(chunks=16, agents=100_000, goals=16, radius=[1.5, 3.0], spacing=2.2, seed=0x5EED0001, agent_seed=0x5EED0001, grid=None,
               desc="1024x1024-tile map (16x16 chunks), 100000 agents/GPU in 16 flocks/GPU (radii 1.5 / 3.0), 16 flow-field goals/GPU"),
    "C3": dict(chunks=16, agents=1_000_000, goals=64, radius=[1.0], spacing=4.05, seed=0x5EED0001, agent_seed=0x5EED0003, grid=None,
               desc="1024x1024-tile map (16x16 chunks), 1000000 agents/GPU of radius 1.0 in 64 flocks/GPU, 64 goals/GPU"),
    "C4": dict(chunks=32, agents=500_000, goals=8, radius=[1.0], spacing=3.9, seed=0x5EED0004, agent_seed=0x5EED0004, grid=8,
               desc="2048x2048-tile map (32x32 chunks), 500000 agents/GPU of radius 1.0 in 8 flocks/GPU (each flock in its own "
                    "cell of an 8x8 grid: constant density), 8 goals/GPU; N = 8 is BASELINE configs[3] (4 M agents, 64 goals)"),
    "C5": dict(chunks=16, agents=62_500, goals=2, radius=[1.0], spacing=3.9, seed=0x5EED0001, agent_seed=0x5EED0005, grid=4,
               blockers=1000,
               desc="1024x1024-tile map, 62500 agents/GPU of radius 1.0 in 2 flocks/GPU (own cells of a 4x4 grid), 2 goals/GPU, "
                    "1000 dynamic obstacles of radius 6 that all move one tile per tick (replicated on every rank); "
                    "N = 8 is BASELINE configs[4] (500 k agents)"),
}
WORKLOAD = "C2"
# module-level mirrors of the selected workload (tests and tools read them)
CHUNKS, AGENTS_PER_GPU, GOALS_PER_GPU, MAP_SEED = 16, 100_000, 16, 0x5EED0001


def set_workload(name):
    """select the synthetic configuration the module-level sizes describe"""
    global WORKLOAD, CHUNKS, AGENTS_PER_GPU, GOALS_PER_GPU, MAP_SEED
    w = WORKLOADS[name]
    WORKLOAD, CHUNKS, AGENTS_PER_GPU, GOALS_PER_GPU, MAP_SEED = name, w["chunks"], w["agents"], w["goals"], w["seed"]


def shard_range(n, rank, world):
    """contiguous, balanced [lo, hi) of n items for `rank` (movement.c:3751-3762 equal-range split)"""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler(threading.Thread):
    """samples nvidia-smi clocks + throttle reasons during the timed region"""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.stop_evt = threading.Event()
        self.samples = []
        self.reasons = set()
        self.sm_max = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0])); self.sm_max = float(out[1])
                for n, v in zip(names, out[2:]):
                    if "Active" in v and "Not" not in v:
                        self.reasons.add(n)
            except Exception:
                pass
            self.stop_evt.wait(0.2)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons)}


def build_workload(pf, world, rank):
    """the selected synthetic workload of this rank: map, whole population (weak scaling: world x per-GPU), goals"""
    synth, capi = pf.synth, pf.capi
    w = WORKLOADS[WORKLOAD]
    p = synth.make_map(CHUNKS, CHUNKS, MAP_SEED, rivers=CHUNKS > 1)
    cost = synth.cost_from_pathable(p, CHUNKS, CHUNKS)
    n_total = AGENTS_PER_GPU * world
    nflocks = GOALS_PER_GPU * world
    radii = np.array([w["radius"][f % len(w["radius"])] for f in range(nflocks)], np.float32)
    cells = (w["grid"], 0) if w["grid"] else None
    a = synth.make_agents(cost, CHUNKS, CHUNKS, n_total, nflocks, w["agent_seed"], radius=radii, spacing=w["spacing"], hz=HZ,
                          cells=cells)
    # field-pool destinations are rank-local: flock f of this rank's goal range -> dest f - g_lo
    g_lo, g_hi = shard_range(nflocks, rank, world)
    dest = np.full(nflocks, -1, np.int32)
    dest[g_lo:g_hi] = np.arange(g_hi - g_lo)
    a["flock_dest_index"] = dest
    rec, fl = capi.pack_agents(a)
    lo, hi = shard_range(n_total, rank, world)
    return dict(pathable=p, cost=cost, agents=a, rec=rec, flocks=fl, lo=lo, hi=hi, g_lo=g_lo, g_hi=g_hi,
                n_total=n_total, nflocks=nflocks)


def neighbour_stats(a, sample=2000, seed=0):
    """k10 / k30: mean number of other agents within 10 / 30 wu (ClearPath cost is cubic in k10)"""
    rng = np.random.default_rng(seed)
    pos = a["pos"]
    idx = rng.integers(0, len(pos), min(sample, len(pos)))
    cell = np.floor(pos / 32.0).astype(np.int64)
    key = cell[:, 0] * 100003 + cell[:, 1]
    order = np.argsort(key, kind="stable"); skey = key[order]
    k10 = k30 = 0
    for i in idx:
        cand = []
        for dx in (-1, 0, 1):
            for dz in (-1, 0, 1):
                k = (cell[i, 0] + dx) * 100003 + (cell[i, 1] + dz)
                l, r = np.searchsorted(skey, k, "left"), np.searchsorted(skey, k, "right")
                cand.append(order[l:r])
        cand = np.concatenate(cand)
        d = np.linalg.norm(pos[cand] - pos[i], axis=1)
        k10 += (d <= 10.0).sum() - 1; k30 += (d <= 30.0).sum() - 1
    return k10 / len(idx), k30 / len(idx)


class Churn:
    """C5: 1 000 circular obstacles of radius 6, a seeded herd around each of this box's flocks, every one moves by one
    tile per tick (N_BlockersDecref at the old position + N_BlockersIncref at the new one, nav.c:4663-4683). The same
    sequence runs on every rank (the map is replicated; in the engine the deltas come from the simulation thread)."""

    def __init__(self, a, nblockers, chunks, seed):
        rng = np.random.default_rng(seed)
        ctr = np.stack([a["pos"][a["flock_of"] == f].mean(axis=0) for f in range(len(a["flock_target"]))])
        self.pos = (ctr[rng.integers(0, len(ctr), nblockers)] + rng.normal(scale=220.0, size=(nblockers, 2))).astype(np.float32)
        self.lim = chunks * 256 - 24
        self.clip()
        self.rng = rng

    def clip(self):
        self.pos[:, 0] = np.clip(self.pos[:, 0], -self.lim, -24); self.pos[:, 1] = np.clip(self.pos[:, 1], 24, self.lim)

    def ops(self, capi, pos, delta):
        o = np.zeros(len(pos), capi.BLOCKER_OP)
        o["x"] = pos[:, 0]; o["z"] = pos[:, 1]; o["range"] = 6.0; o["flags"] = capi.FLAG_MOVABLE; o["delta"] = delta
        return o

    def place(self, nav, capi):
        nav.blockers_batch(self.ops(capi, self.pos, +1))

    def move(self, nav, capi):
        old = self.pos.copy()
        self.pos += self.rng.integers(-1, 2, size=self.pos.shape).astype(np.float32) * 4.0
        self.clip()
        both = np.empty(2 * len(old), capi.BLOCKER_OP)
        both[0::2] = self.ops(capi, old, -1); both[1::2] = self.ops(capi, self.pos, +1)     # decref old, incref new, per obstacle
        nav.blockers_batch(both)


def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("bench.py --gpus %d must be launched with torch.distributed.run --nproc-per-node %d" % (args.gpus, args.gpus))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    pf = importlib.import_module("permafrost-engine_b200")
    capi = pf.capi
    wl = WORKLOADS[WORKLOAD]
    W = build_workload(pf, world, rank)
    nav = capi.Nav(local_rank)
    nav.map_create(CHUNKS, CHUNKS, 1)
    nav.map_upload_layer(0, W["cost"])
    nav.map_build_nav(0)
    churn = None
    if wl.get("blockers"):
        nav.route_build(0)                      # edge states follow the blockers (N_Update, nav.c:2119)
        churn = Churn(W["agents"], wl["blockers"], CHUNKS, MAP_SEED + 77)
        churn.place(nav, capi)
        nav.map_commit()
    ngoals = W["g_hi"] - W["g_lo"]
    nav.pool_create(ngoals, ngoals * CHUNKS * CHUNKS)
    if os.environ.get("PFNAV_COHESION_MODE"):          # A/B hook: 1 always the windowed cohesion pass, 2 always the member list
        nav.set_cohesion_mode(int(os.environ["PFNAV_COHESION_MODE"]))
    if os.environ.get("PFNAV_TWO_PHASE"):              # A/B and profiling hook: 0 single pass, 2 always split
        nav.set_two_phase(int(os.environ["PFNAV_TWO_PHASE"]))
    goals = [tuple(int(v) for v in W["agents"]["flock_target_tile"][f]) for f in range(W["g_lo"], W["g_hi"])]

    n_total, lo, hi = W["n_total"], W["lo"], W["hi"]
    if world > 1:
        # multi-GPU lives in the library: torch.distributed only hands the NCCL id around and reduces the timings
        ids = [capi.mgpu_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        nav.mgpu_init(rank, world, ids[0])
        assert capi.mgpu_shard_range(n_total, rank, world) == (lo, hi)
    # pinned host copies for the e2e leg: this rank's own records only
    nwork = hi - lo
    rec_pinned = torch.empty(max(nwork, 1) * capi.AGENT.itemsize, dtype=torch.uint8).pin_memory()
    rec_np = rec_pinned.numpy().view(capi.AGENT)[:nwork]
    rec_np[:] = W["rec"][lo:hi]
    work = np.arange(lo, hi, dtype=np.uint32)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    sp = stream.cuda_stream
    assert sp != 0

    def upload(flags=0):
        nav.agents_upload_shard(rec_np, lo, hi, n_total, W["flocks"], HZ, flags)

    upload()
    if world > 1:
        nav.mgpu_gather(sp)
    nav.agents_set_work(work)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")        # > 126 MB L2
    goal_dests = np.arange(len(goals), dtype=np.int32)
    goal_targets = np.array(goals, np.int32)
    nfl = [0, 0]

    def fields_phase(mode):
        if churn is not None:
            churn.move(nav, capi)               # 2 x 1000 refcount updates on the host mirrors ...
            nav.map_commit()                    # ... N_Update: islands, edge states, pool invalidation, chunk upload
            nfl[:] = nav.pool_request_goals(goal_dests, goal_targets, 0, sp, capi.REQUEST_MISSING_ONLY)    # only what N_Update invalidated
        elif mode == "cold":
            nav.pool_clear()
            nfl[:] = nav.pool_request_goals(goal_dests, goal_targets, 0, sp)
        elif mode == "resident":
            nfl[:] = nav.pool_request_goals(goal_dests, goal_targets, 0, sp)

    def gather_phase():
        if world > 1:
            nav.mgpu_gather(sp)                  # ONE all-gather of 24-byte records + index rebuild, in the library
        else:
            nav.agents_rebuild_index(sp)

    def step_resident(mode):
        fields_phase(mode)
        gather_phase()
        nav.agents_tick(capi.TICK_VDES_FROM_POOL, sp)

    def step_e2e(mode):
        fields_phase(mode)                                    # asynchronous; the LOS chains overlap the upload below
        upload(capi.UPLOAD_SAME_FLOCKS)                       # H2D of this rank's records from pinned memory (+ index at N = 1)
        if world > 1:
            nav.mgpu_gather(sp)
        nav.agents_tick(capi.TICK_VDES_FROM_POOL, sp)
        return nav.agents_read_velocities(nwork)              # D2H of the result

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(stepfn, steps):
        """K steps, CUDA events around every step on the one stream, L2 flushed between steps; -> (device ms, launches, wall)"""
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        launches0 = nav.launch_count()
        barrier()
        t0 = time.perf_counter()
        for k in range(steps):
            flush.zero_()                                # L2 flush between timed iterations (outside the event pair)
            ev[k][0].record(stream)
            stepfn()
            ev[k][1].record(stream)
        barrier()
        wall = time.perf_counter() - t0
        ms = float(sum(a.elapsed_time(b) for a, b in ev))
        launches = nav.launch_count() - launches0
        if world > 1:
            t = torch.tensor([ms, float(launches)], device="cuda")
            tm = t.clone(); dist.all_reduce(tm, op=dist.ReduceOp.MAX); dist.all_reduce(t)
            ms, launches = float(tm[0].item()), int(t[1].item())
        return ms, launches, wall

    headline = "cold"
    # ---- warm-up ----
    for _ in range(max(args.warmup, 3)):
        step_resident(headline)
    torch.cuda.synchronize()

    # ---- timed: K steps, device-resident, COLD (the headline) ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    dev_ms, launches, t_wall = timed(lambda: step_resident(headline), args.steps)
    nf, nl = nfl
    extra = {}
    if churn is None:
        step_resident("resident")
        extra["resident"] = timed(lambda: step_resident("resident"), args.steps)[0]
        extra["warm"] = timed(lambda: step_resident("warm"), args.steps)[0]

    if os.environ.get("PF_BENCH_DEBUG"):
        nav.profile_enable(True); nav.profile_read()
        for _ in range(3):
            flush.zero_(); step_resident(headline)
        torch.cuda.synchronize()
        print("debug: in-step group times over 3 steps (ms, launches): %s" % nav.profile_read(), file=sys.stderr)
        nav.profile_enable(False)

    # ---- per-phase device times, each phase alone between synchronisations ----
    def phase_time(fn, iters=5):
        ms, profs = [], []
        for _ in range(iters):
            flush.zero_(); torch.cuda.synchronize()
            a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            a_.record(stream); fn(); b_.record(stream); torch.cuda.synchronize()
            ms.append((a_.elapsed_time(b_), (time.perf_counter() - t0) * 1e3))
            profs.append(nav.profile_read())
        med = {k: (float(np.median([p_[k][0] for p_ in profs])), profs[0][k][1]) for k in profs[0]}
        return float(np.median([m[0] for m in ms])), float(np.median([m[1] for m in ms])), med
    nav.profile_enable(True)
    nav.profile_read()
    fields_cold_dev_ms, fields_cold_wall_ms, prof_f = phase_time(lambda: (fields_phase(headline), nav.fields_join(sp)))
    _, _, prof_i = phase_time(gather_phase)
    tick_alone_ms, _, prof_t = phase_time(lambda: nav.agents_tick(capi.TICK_VDES_FROM_POOL, sp))
    nav.profile_enable(False)
    prof = {"flow": prof_f["flow"], "los": prof_f["los"], "index": prof_i["index"], "vdes": prof_t["vdes"],
            "cohesion": prof_t["cohesion"], "velocity": prof_t["velocity"]}
    vpref, vdes, los = nav.agents_read_debug(nwork)
    frac_no_dir = float((np.abs(vdes).sum(axis=1) == 0).mean())
    if frac_no_dir > 0.01 and churn is None:
        raise SystemExit("bench.py: %.2f%% of the agents found no flow direction in the field pool -- the step did not do "
                         "the work it claims" % (100 * frac_no_dir))

    # ---- timed: e2e through the host-buffer ABI (cold fields, like the headline) ----
    for _ in range(2):
        step_e2e(headline)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()
        step_e2e(headline)
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); e2e_s = float(t.item())

    # ---- device-resident multi-tick run: positions advance, fields warm ----
    moving = None
    if churn is None and not args.no_moving:
        nav.route_build(0)
        ms0 = np.zeros(nwork, capi.MOVESTATE)
        ms0["next_pos"][:, 0] = rec_np["pos"][:, 0]; ms0["next_pos"][:, 2] = rec_np["pos"][:, 1]
        ms0["step"] = 1.0; ms0["next_rot"][:, 3] = 1.0; ms0["combat_facing"][:, 3] = 1.0
        ms0["vel_hist"] = np.repeat(rec_np["velocity"][:, None, :], 14, axis=1)
        upload(); gather_phase()
        nav.agents_upload_movestate(ms0)
        nav.agents_set_work(work)

        def step_moving():
            nav.agents_tick(capi.TICK_VDES_FROM_POOL, sp)
            nav.agents_compute_updates(sp)
            nav.agents_apply_updates(sp)
            gather_phase()
        nav.pool_request_goals(goal_dests, goal_targets, 0, sp)
        step_moving()
        nav.clearpath_stats()                                       # reset
        ms_m, _, _ = timed(step_moving, args.steps)
        cps = nav.clearpath_stats()
        nav.profile_enable(True); nav.profile_read()
        for _ in range(3):
            step_moving()
        torch.cuda.synchronize()
        pm = nav.profile_read(); nav.profile_enable(False)
        st, _ = nav.agents_read_state(nwork, movestate=False)
        moved = float(np.linalg.norm(st["pos"] - rec_np["pos"], axis=1).mean())
        moving = {"value": nwork * world * args.steps / (ms_m / 1e3), "ms_per_step": ms_m / args.steps,
                  "mean_displacement_wu": moved, "still_moving": float(((st["state"] != 2) & (st["state"] != 4)).mean()),
                  "phase_ms_per_step": {k: v[0] / 3 for k, v in pm.items() if v[1]},
                  "clearpath_per_step": {k: v / args.steps for k, v in cps.items()},
                  "note": "tick + entity_compute_update + entity_apply_update + gather/index per step, positions advance on the device"}
    if rank == 0:
        sampler.stop_evt.set(); sampler.join(2)

    total_agents = nwork * world
    ms_per_step = dev_ms / args.steps
    value = total_agents * args.steps / (dev_ms / 1e3)
    peaks, peak_src = measured_peaks()
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    groups = {k: v[0] for k, v in prof.items()}              # ms per step (each group alone)
    dom = max(groups, key=groups.get)
    per_launch = {"velocity": nwork * ALG_BYTES_PER_AGENT, "cohesion": nwork * ALG_BYTES_PER_AGENT,
                  "flow": nf * ALG_BYTES_PER_FLOW_FIELD, "los": nl * ALG_BYTES_PER_LOS_FIELD,
                  "index": n_total * 24 * 2, "vdes": nwork * 64}
    dom_ms = groups[dom]
    achieved = per_launch[dom] / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    traffic_tab = {}
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic_tab = json.load(f)
    except Exception:
        pass
    kname = {"velocity": "k_agent_velocity", "cohesion": "k_cohesion", "flow": "k_flow_unit", "los": "k_los_b",
             "index": "k_cell_*", "vdes": "k_desired_velocity"}
    sm_count = torch.cuda.get_device_properties(local_rank).multi_processor_count
    issue_peak = sm_count * ISSUE_PEAK_PER_SM_CLK * float(peaks.get("sm_max_mhz", 1965.0)) * 1e6     # warp-instructions / s
    issue = {}
    for g_ in ("velocity", "los"):
        per_unit = traffic_tab.get("warp_inst_per_unit", {}).get(WORKLOAD, {}).get(g_)
        units = nwork if g_ == "velocity" else nl
        if per_unit and groups[g_] > 0:
            ach = per_unit * units / (groups[g_] * 1e-3)
            issue[kname[g_]] = {"warp_inst_per_unit": per_unit, "achieved_ginst_s": ach / 1e9, "peak_ginst_s": issue_peak / 1e9,
                                "frac": ach / issue_peak, "source": "ncu sm__inst_executed.sum per launch (profiles/), "
                                "divided by this run's measured duration"}
    fields_dev = groups["flow"] + groups["los"]
    result = {
        "metric": "agent_updates_per_sec", "value": value, "unit": "agent-updates/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD + ": " + wl["desc"] + ", hz=20",
                   "agents_total": total_agents, "goals_total": GOALS_PER_GPU * world, "map_seed": hex(MAP_SEED),
                   "headline": "COLD step: pool cleared, every goal re-planned on the host and all its fields rebuilt inside the "
                               "timed region, then gather + tick" if churn is None else
                               "churn step: 1000 obstacles moved + N_Update + invalidated fields rebuilt + tick, all inside the timed region",
                   "parallelism": ("agents + goals sharded x%d inside libpfnav.so (pfnav_mgpu_*): 1 NCCL all-gather of 24-B "
                                   "records per tick, map replicated" % world) if world > 1 else "single GPU",
                   "l2": "256 MiB memset between timed steps (outside the per-step event pairs); working set < L2",
                   "agents_without_flow_direction": frac_no_dir},
        "flow_fields_per_sec": (nf + nl) * world / (fields_cold_wall_ms * 1e-3) if fields_cold_wall_ms > 0 else None,
        "flow_fields_per_sec_kernels_only": (nf + nl) * world / (fields_dev * 1e-3) if fields_dev > 0 else None,
        "fields_per_step": {"flow": nf * world, "los": nl * world,
                            "note": "flow_fields_per_sec = fields of one cold request batch / its wall time incl. host planning and "
                                    "upload; *_kernels_only = the same fields / summed kernel time"},
        "phase_ms_per_step": dict(groups, fields_cold_wall=fields_cold_wall_ms, fields_cold_device=fields_cold_dev_ms, tick_alone=tick_alone_ms),
        "wall_s": t_wall,
        "gpu_launches": launches,
        "e2e": {"value": total_agents * args.steps / e2e_s, "unit": "agent-updates/s",
                "h2d_bytes_per_step": int(nwork * capi.AGENT.itemsize),
                "d2h_bytes_per_step": int(nwork * 8),
                "note": "per rank: own 64-B entity records up, own velocities down; cold fields like the headline"},
        "roofline": {"bound": "hbm", "kernel": kname[dom],
                     "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm, "traffic": traffic_tab.get(dom),
                     "peak_source": peak_src + " (MEASURED_PEAKS.json hbm_gbs)" if peak_src == "measured" else "fallback 6650 GB/s",
                     "algorithmic_bytes_per_launch": per_launch[dom], "ms_per_launch": dom_ms,
                     "note": "latency/ALU-bound by construction: 360 B of compulsory traffic per agent update (SURVEY.md 8d); "
                             "the issue-slot figures are the meaningful utilisation of these kernels",
                     "issue_slots": issue},
        "clocks": sampler.summary() if rank == 0 else None,
    }
    if "warm" in extra:
        result["value_warm"] = total_agents * args.steps / (extra["warm"] / 1e3)
        result["value_resident_plan"] = total_agents * args.steps / (extra["resident"] / 1e3)
        result["ms_per_step_modes"] = {"cold": ms_per_step, "resident_plan": extra["resident"] / args.steps, "warm": extra["warm"] / args.steps}
    if moving:
        result["value_moving"] = moving
    if rank == 0:
        k10, k30 = neighbour_stats(W["agents"])
        result["config"]["k10_mean"] = k10; result["config"]["k30_mean"] = k30
        if world == 1 and not args.no_cpu_baseline:
            result["cpu_baseline"] = cpu_baseline(pf, budget_s=args.cpu_budget)
        print(json.dumps(result), flush=True)
    if world > 1:
        nav.mgpu_finalize()
    nav.close()
    if world > 1:
        dist.destroy_process_group()


def _ref_or_port():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pfref
    if pfref.available():
        return "reference", pfref
    if os.path.isdir("/root/reference/src"):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"], check=False)
        if pfref.available():
            return "reference", pfref
    import pforacle
    return "port", pforacle


_CPU_SETUP = {}


def cpu_setup(pf):
    """the CPU arm's bounded sample of the selected workload, built once: a map of the same generator small enough for the
    reference's 665 KB-per-chunk layout to be built in seconds, flocks of the workload's own recipe (size, radii, spacing)
    with the WHOLE sample population present as neighbours, one N_RequestPath per flock from its centre."""
    if WORKLOAD in _CPU_SETUP:
        return _CPU_SETUP[WORKLOAD]
    synth, capi = pf.synth, pf.capi
    kind, mod = _ref_or_port()
    wl = WORKLOADS[WORKLOAD]
    per_flock = wl["agents"] // wl["goals"]
    if WORKLOAD == "C1":
        cw, nflocks = 1, 1
    else:
        cw = 8 if kind == "reference" else 4
        nflocks = max(2, min(wl["goals"], (70_000 if kind == "reference" else 8_000) // max(per_flock, 1) + 1))
    p = synth.make_map(cw, cw, MAP_SEED, rivers=cw > 1)
    cost = synth.cost_from_pathable(p, cw, cw)
    radii = np.array([wl["radius"][f % len(wl["radius"])] for f in range(nflocks)], np.float32)
    grid = None
    if wl["grid"]:
        side = int(np.ceil(np.sqrt(nflocks)))
        grid = (max(side, int(round(cw * wl["grid"] / wl["chunks"]))), 0)
    n = per_flock * nflocks
    a = synth.make_agents(cost, cw, cw, n, nflocks, wl["agent_seed"], radius=radii, spacing=wl["spacing"], hz=HZ, cells=grid)
    S = dict(kind=kind, mod=mod, cw=cw, p=p, cost=cost, a=a, n=n, nflocks=nflocks)
    rng = np.random.default_rng(1)
    tiles = synth.random_passable_tiles(cost, 1024, rng)
    S["field_reqs"] = np.array([[t[0] // cw, t[0] % cw, t[1], t[2]] for t in tiles], np.int32)
    if kind == "reference":
        ref = mod.RefMap(cw, cw, p)
        dest = np.zeros(nflocks, np.uint32)
        for f in range(nflocks):
            pts = a["pos"][a["flock_of"] == f]
            src = pts[np.argmin(np.linalg.norm(pts - pts.mean(axis=0), axis=1))]
            ok, did = ref.request_path((float(src[0]), float(src[1])), (float(a["flock_target"][f][0]), float(a["flock_target"][f][1])))
            dest[f] = did
        ref.agents_set(a["pos"], a["prev_pos"], a["vel"], a["radius"], a["max_speed"], a["state"], a["flags"],
                       a["flock_of"], a["flock_target"], dest, hz=HZ)
        S["ref"] = ref
    else:
        om = mod.OracleMap(cw, cw, cost)
        aa = dict(a)
        d = a["flock_target"][a["flock_of"]] - a["pos"]
        aa["vdes"] = (d / np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-6)).astype(np.float32)
        rec, fl = capi.pack_agents(aa)
        S["om"] = om; S["world"] = mod.OracleWorld(om, rec, fl, HZ)
    _CPU_SETUP[WORKLOAD] = S
    return S


def cpu_tick(S, work, threads):
    """one CPU tick over `work`: A2 = compute_los_state + compute_desired_velocity (movement.c:4129-4180, serial on the
    navigation fiber like in the engine, field cache warm after the first call) + A1 = move_velocity_work over the
    reference's equal-range fork-join (movement.c:3751-3762) on `threads` threads. -> seconds (A2, A1)"""
    a = S["a"]
    if S["kind"] == "reference":
        ref = S["ref"]
        ref.work_set(work, np.zeros((len(work), 2), np.float32), np.zeros(len(work), np.uint8), a["speed"][work])
        t0 = time.perf_counter()
        ref.desired_from_cache()
        t_a2 = time.perf_counter() - t0
        _, t_a1 = ref.velocity_work(threads)
        return t_a2, t_a1
    t0 = time.perf_counter(); S["world"].velocity_work(work); return 0.0, time.perf_counter() - t0


def cpu_sample(pf, budget_s):
    """Bounded CPU sample of the selected workload, `budget_s` seconds of CPU wall: repeated whole-sample ticks on all
    host cores (median of >= 5), a 1-thread figure on a sub-sample, flow + LOS field batches. Returns a dict."""
    S = cpu_setup(pf)
    cores = os.cpu_count() or 1
    threads = cores if S["kind"] == "reference" else 1
    n = S["n"]
    work_all = np.arange(n, dtype=np.uint32)
    t_start = time.perf_counter()
    cpu_tick(S, work_all[:: max(1, n // 4096)], threads)                   # warm the field cache / page in
    runs = []
    while len(runs) < 5 or (time.perf_counter() - t_start < 0.6 * budget_s and len(runs) < 25):
        runs.append(cpu_tick(S, work_all, threads))
    tick_s = np.array([r[0] + r[1] for r in runs])
    rng = np.random.default_rng(2)
    sub = np.sort(rng.choice(n, min(n, 6000), replace=False)).astype(np.uint32)
    one = [cpu_tick(S, sub, 1) for _ in range(2)]
    one_s = min(r[0] + r[1] for r in one)
    reqs = S["field_reqs"]
    if S["kind"] == "reference":
        f_secs, _ = S["ref"].fields_mt(0, reqs, threads)
        l_secs, _ = S["ref"].fields_mt(1, reqs, threads)
        f1, _ = S["ref"].fields_mt(0, reqs[:128], 1); l1, _ = S["ref"].fields_mt(1, reqs[:128], 1)
        fields_per_s, fields_per_s_1t = 2 * len(reqs) / (f_secs + l_secs), 2 * 128 / (f1 + l1)
    else:
        capi = pf.capi
        fr = np.concatenate([capi.tile_req((int(q[0]), int(q[1])), (int(q[2]), int(q[3]))) for q in reqs])
        lr = np.concatenate([capi.los_req((int(q[0]), int(q[1])), (int(q[0]), int(q[1]), int(q[2]), int(q[3]))) for q in reqs])
        t0 = time.perf_counter(); S["om"].flow_fields_update(fr); S["om"].los_fields_create(lr)
        fields_per_s = fields_per_s_1t = 2 * len(reqs) / (time.perf_counter() - t0)
    med = float(np.median(tick_s))
    a2 = float(np.median([r[0] for r in runs]))
    desc = ("%d agents = %d whole flocks of the %s recipe (%d each, whole sample present as neighbours) on a %dx%d-chunk map of "
            "the same generator; per tick A2 (compute_los_state + compute_desired_velocity out of the warm field cache, serial: "
            "median %.3f s) + A1 (move_velocity_work, the reference's equal-range split over %d threads); median of %d ticks "
            "(min %.3f / max %.3f s); 1-thread figure on a %d-agent sub-sample; 1024 destination-chunk flow + 1024 LOS fields; "
            "%.1f s of CPU wall" % (n, S["nflocks"], WORKLOAD, n // S["nflocks"], S["cw"], S["cw"], a2, threads, len(runs),
                                   float(tick_s.min()), float(tick_s.max()), len(sub), time.perf_counter() - t_start))
    return dict(value=n / med, value_1thread=len(sub) / one_s, cores=threads, kind=S["kind"], sample=desc,
                flow_fields_per_sec=fields_per_s, flow_fields_per_sec_1thread=fields_per_s_1t, ticks=len(runs),
                spread=float(tick_s.max() / tick_s.min()))


def cpu_baseline(pf, budget_s=20.0):
    r = cpu_sample(pf, budget_s)
    return {"value": r["value"], "unit": "agent-updates/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"],
            "value_1thread": r["value_1thread"], "flow_fields_per_sec": r["flow_fields_per_sec"],
            "flow_fields_per_sec_1thread": r["flow_fields_per_sec_1thread"], "run_to_run_spread": r["spread"],
            "host_cpus": os.cpu_count()}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pf = importlib.import_module("permafrost-engine_b200")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    vals, fvals, last = [], [], None
    t0 = time.perf_counter()
    per_step = max(2.0, min(args.cpu_budget, 200.0 / max(1, args.warmup + args.steps)))
    for k in range(args.warmup + args.steps):
        last = cpu_sample(pf, per_step)
        if k >= args.warmup:
            vals.append(last["value"]); fvals.append(last["flow_fields_per_sec"])
        if time.perf_counter() - t0 > 240 and len(vals) >= 1:
            break
    value = float(np.median(vals))
    wl = WORKLOADS[WORKLOAD]
    out = {
        "impl": "reference", "metric": "agent_updates_per_sec", "value": value, "unit": "agent-updates/s",
        "n_gpus": args.gpus, "steps": len(vals), "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD + ": " + wl["desc"] + ", hz=20",
                   "note": "CPU arm: throughput of the reference's own code (A2 + A1 per tick) on a bounded sample of that workload; "
                           "it does not scale with --gpus (rank 0 only, world=%d)" % world},
        "flow_fields_per_sec": float(np.median(fvals)),
        "cpu_baseline": {"value": value, "unit": "agent-updates/s", "cores": last["cores"], "kind": last["kind"], "sample": last["sample"],
                         "value_1thread": last["value_1thread"], "host_cpus": os.cpu_count(),
                         "run_to_run_spread_over_steps": float(max(vals) / min(vals))},
        "e2e": {"value": value, "unit": "agent-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


def main():
    # stdout carries exactly one JSON line: NCCL's own banner ("NCCL version ...", printed when the box sets NCCL_DEBUG) goes to a file
    os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/pfnav_bench_nccl.%h.%p.log")
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-moving", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=20.0)
    ap.add_argument("--workload", default="auto", choices=["auto"] + sorted(WORKLOADS),
                    help="auto: C2 (BASELINE.json configs[1]) on one GPU, C4 (configs[3] at 8 GPUs; 500 k agents + 8 goals per GPU at "
                         "constant density) on several")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    set_workload(("C2" if max(world, args.gpus) == 1 else "C4") if args.workload == "auto" else args.workload)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
