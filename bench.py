#!/usr/bin/env python3
"""bench.py -- the driver's measurement contract for the navigation / crowd-movement hot path.

    python bench.py --gpus N --steps K --warmup W            (N=1; N>1 via torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W

A "step" is one movement tick over one synthetic batch (BASELINE.json configs[1]: 1024x1024-tile map
= 16x16 chunks, 100k agents in 16 flocks, 16 concurrent flow-field goals, per GPU):
   1. the 16 goals' field sets are (re)built into the device field pool
      (pfnav_pool_request_goal: flow waves + LOS chain, every chunk connected to the goal),
   2. (N>1) one NCCL all-gather of the 24-byte neighbour records, then the spatial index rebuild,
   3. the agent tick: desired-velocity gather from the pool, cohesion, boids + ClearPath velocity.
`value` = agent updates per second over the whole step with everything resident in HBM;
`flow_fields_per_sec` is reported beside it from the same steps. `e2e` repeats the measurement
through the host-buffer C ABI (agent records H2D from pinned memory, velocities D2H, every step).

Weak scaling: every rank owns 100k agents / 16 goals of its own; the population grows with N, every
rank indexes the whole population (all-gather), the map is replicated.

--impl reference times the reference's own CPU implementation (oracle/_ref, the unmodified sources
compiled by oracle/Makefile; falls back to the C port when that is absent) on the host cores, on a
bounded sample of the same workload.
"""
import argparse
import ctypes
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

MAP_SEED = 0x5EED0001          # 0x5EED0000 + config#
CHUNKS = 16
AGENTS_PER_GPU = 100_000
GOALS_PER_GPU = 16
WORKLOAD = "C2"
C3_MODE = False
HZ = 20
ALG_BYTES_PER_AGENT = 360      # SURVEY.md 8d "Canonical figure"
ALG_BYTES_PER_FLOW_FIELD = 24_704   # TARGET_PORTAL: cost 4096 + blockers 8192 + islands 8192+128 + dirs 4096
ALG_BYTES_PER_LOS_FIELD = 16_512


def set_workload(name):
    """select the synthetic configuration the module-level sizes describe"""
    global AGENTS_PER_GPU, GOALS_PER_GPU, WORKLOAD, C3_MODE, CHUNKS
    if name == "C2":
        AGENTS_PER_GPU, GOALS_PER_GPU, WORKLOAD, C3_MODE, CHUNKS = 100_000, 16, "C2", False, 16
    elif name == "C3":
        AGENTS_PER_GPU, GOALS_PER_GPU, WORKLOAD, C3_MODE, CHUNKS = 1_000_000, 64, "C3", True, 16
    else:
        raise ValueError(name)


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
        return {"hbm_gbs": 6650.0}, "fallback"


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


class CudaArrayView:
    """exposes a raw device pointer to torch through __cuda_array_interface__ (zero copy)"""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def build_workload(pf, world, rank):
    """the synthetic C2 workload of this rank: map, whole population (weak scaling: world x 100k), goals"""
    synth, capi = pf.synth, pf.capi
    p = synth.make_map(CHUNKS, CHUNKS, MAP_SEED)
    cost = synth.cost_from_pathable(p, CHUNKS, CHUNKS)
    n_total = AGENTS_PER_GPU * world
    nflocks = GOALS_PER_GPU * world
    # radii in {1.5, 3.0}, spawn discs with spacing >= 2.2 r
    radii = np.where(np.arange(nflocks) % 2 == 0, 1.5, 3.0).astype(np.float32)
    if C3_MODE:
        # radius 1.0, 64 flocks. 1 M agents on 14 M wu^2 of passable ground cannot be sparser than
        # ~0.07 agents/wu^2 (k10 ~ 22), so the spawn discs are sized to tile the map (spacing 4.05 r)
        a = synth.make_agents(cost, CHUNKS, CHUNKS, n_total, nflocks, MAP_SEED + 2, radius=1.0, spacing=4.05, hz=HZ)
    else:
        a = synth.make_agents(cost, CHUNKS, CHUNKS, n_total, nflocks, MAP_SEED, radius=radii, spacing=2.2, hz=HZ)
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
    # grid-bucketed count to stay O(n)
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


def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus %d must be launched with torch.distributed.run --nproc-per-node %d" % (args.gpus, args.gpus))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    pf = importlib.import_module("permafrost-engine_b200")
    capi = pf.capi
    W = build_workload(pf, world, rank)
    nav = capi.Nav(local_rank)
    nav.map_create(CHUNKS, CHUNKS, 1)
    nav.map_upload_layer(0, W["cost"])
    nav.map_build_nav(0)
    ngoals = W["g_hi"] - W["g_lo"]
    nav.pool_create(ngoals, ngoals * CHUNKS * CHUNKS)
    if os.environ.get("PFNAV_TWO_PHASE"):              # A/B and profiling hook: 0 single pass, 2 always split
        nav.set_two_phase(int(os.environ["PFNAV_TWO_PHASE"]))
    goals = [tuple(int(v) for v in W["agents"]["flock_target_tile"][f]) for f in range(W["g_lo"], W["g_hi"])]

    # pinned host copies for the e2e leg
    n_total = W["n_total"]
    rec_pinned = torch.empty(n_total * capi.AGENT.itemsize, dtype=torch.uint8).pin_memory()
    rec_np = rec_pinned.numpy().view(capi.AGENT)
    rec_np[:] = W["rec"]
    work = np.arange(W["lo"], W["hi"], dtype=np.uint32)
    nwork = len(work)
    vel_pinned = torch.empty(nwork * 8, dtype=torch.uint8).pin_memory()
    vel_np = vel_pinned.numpy().view(np.float32).reshape(nwork, 2)

    nav.agents_upload(rec_np, W["flocks"], HZ)
    nav.agents_set_work(work)
    # one explicit (non-default) stream carries the whole step: torch ops, NCCL and the library calls
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    sp = stream.cuda_stream
    assert sp != 0
    d_rec_ptr, _, _ = nav.agents_device_ptrs()
    rec_view = torch.as_tensor(CudaArrayView(d_rec_ptr, n_total * 24), device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")        # > 126 MB L2

    goal_dests = np.arange(len(goals), dtype=np.int32)
    goal_targets = np.array(goals, np.int32)

    def fields_phase():
        return nav.pool_request_goals(goal_dests, goal_targets, 0, sp)

    def gather_phase():
        if world > 1:
            lo, hi = W["lo"], W["hi"]
            counts = [shard_range(n_total, r, world) for r in range(world)]
            if all(c[1] - c[0] == hi - lo for c in counts):
                dist.all_gather_into_tensor(rec_view, rec_view[lo * 24:hi * 24].clone())
            else:
                pad = max(c[1] - c[0] for c in counts) * 24
                buf = torch.zeros(pad, dtype=torch.uint8, device="cuda"); buf[:(hi - lo) * 24] = rec_view[lo * 24:hi * 24]
                outs = [torch.empty(pad, dtype=torch.uint8, device="cuda") for _ in range(world)]
                dist.all_gather(outs, buf)
                for r, c in enumerate(counts):
                    rec_view[c[0] * 24:c[1] * 24] = outs[r][:(c[1] - c[0]) * 24]
        nav.agents_rebuild_index(sp)

    def step_resident():
        nfl = fields_phase()
        gather_phase()
        nav.agents_tick(capi.TICK_VDES_FROM_POOL, sp)
        return nfl

    def step_e2e():
        fields_phase()                                        # asynchronous; the LOS chains overlap the upload below
        nav.agents_upload(rec_np, W["flocks"], HZ)            # H2D of the whole snapshot from pinned memory
        nav.agents_set_work(work)
        nav.agents_tick(capi.TICK_VDES_FROM_POOL, sp)
        v = nav.agents_read_velocities(nwork)                 # D2H of the result
        return v

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up ----
    for _ in range(max(args.warmup, 3)):
        nfl = step_resident()
    torch.cuda.synchronize()
    launches0 = nav.launch_count()

    # ---- timed: K steps, device-resident ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()                                # L2 flush between timed iterations (not inside the event pair)
        ev[k][0].record(stream)
        nfl = step_resident()
        ev[k][1].record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = nav.launch_count() - launches0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    dev_ms = float(sum(step_ms))
    if world > 1:
        t = torch.tensor([dev_ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); dev_ms = float(t.item())
        tl = torch.tensor([float(launches)], device="cuda"); dist.all_reduce(tl); launches = int(tl.item())
    if os.environ.get("PF_BENCH_DEBUG"):
        # kernel-group times INSIDE overlapped steps (library events on each group's own stream)
        nav.profile_enable(True); nav.profile_read()
        for _ in range(3):
            flush.zero_(); step_resident()
        torch.cuda.synchronize()
        pr = nav.profile_read(); nav.profile_enable(False)
        print("debug: in-step group times over 3 steps (ms, launches): %s" % pr, file=sys.stderr)
    # ---- per-phase device times, each phase alone between synchronisations (kernel time without the
    #      host-side gaps of the full step); used for the roofline line and flow_fields_per_sec ----
    def phase_time(fn, iters=5):
        """median over `iters` isolated runs of (CUDA-event time of fn, per-kernel-group times from the library)"""
        ms, profs = [], []
        for _ in range(iters):
            flush.zero_(); torch.cuda.synchronize()
            a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a_.record(stream); fn(); b_.record(stream); torch.cuda.synchronize()
            ms.append(a_.elapsed_time(b_))
            profs.append(nav.profile_read())
        med = {k: (float(np.median([p_[k][0] for p_ in profs])), profs[0][k][1]) for k in profs[0]}
        return float(np.median(ms)), med
    nav.profile_enable(True)
    nav.profile_read()
    _, prof_f = phase_time(lambda: (fields_phase(), nav.fields_join(sp)))
    _, prof_i = phase_time(lambda: nav.agents_rebuild_index(sp))
    tick_alone_ms, prof_t = phase_time(lambda: nav.agents_tick(capi.TICK_VDES_FROM_POOL, sp))
    if os.environ.get("PF_BENCH_DEBUG"):
        print("debug: tick alone %.3f ms, profile %s" % (tick_alone_ms, prof_t), file=sys.stderr)
    nav.profile_enable(False)
    prof = {"flow": prof_f["flow"], "los": prof_f["los"], "index": prof_i["index"], "vdes": prof_t["vdes"],
            "cohesion": prof_t["cohesion"], "velocity": prof_t["velocity"]}
    prof_iters = 1
    # miss counter sanity: every agent must have found its field in the pool
    vpref, vdes, los = nav.agents_read_debug(nwork)
    frac_no_dir = float((np.abs(vdes).sum(axis=1) == 0).mean())
    if frac_no_dir > 0.01:
        raise SystemExit("bench.py: %.2f%% of the agents found no flow direction in the field pool -- the step did not do "
                         "the work it claims" % (100 * frac_no_dir))

    # ---- timed: e2e through the host-buffer ABI ----
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()
        v = step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    if rank == 0:
        sampler.stop_evt.set(); sampler.join(2)
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); e2e_s = float(t.item())

    total_agents = nwork * world
    ms_per_step = dev_ms / args.steps
    value = total_agents * args.steps / (dev_ms / 1e3)
    nf, nl = nfl
    peaks, peak_src = measured_peaks()
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    # dominant kernel by measured device time
    groups = {k: v[0] / prof_iters * args.steps for k, v in prof.items()}      # normalised to `steps` like the rest
    dom = max(groups, key=groups.get)
    per_launch = {"velocity": nwork * ALG_BYTES_PER_AGENT, "cohesion": nwork * ALG_BYTES_PER_AGENT,
                  "flow": nf * ALG_BYTES_PER_FLOW_FIELD, "los": nl * ALG_BYTES_PER_LOS_FIELD,
                  "index": n_total * 24 * 2, "vdes": nwork * 64}
    dom_ms = groups[dom] / args.steps                          # per step == per launch group
    achieved = per_launch[dom] / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get(dom)
    except Exception:
        pass
    result = {
        "metric": "agent_updates_per_sec", "value": value, "unit": "agent-updates/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD + ": 1024x1024-tile map (16x16 chunks), %d agents/GPU in %d flocks/GPU, %d flow-field goals/GPU "
                               "(every chunk connected to each goal), hz=20" % (AGENTS_PER_GPU, GOALS_PER_GPU, GOALS_PER_GPU),
                   "agents_total": total_agents, "goals_total": GOALS_PER_GPU * world, "map_seed": hex(MAP_SEED),
                   "parallelism": "agents+goals sharded x%d, 1 all-gather of 24-B records/tick" % world if world > 1 else "single GPU",
                   "l2": "256 MiB memset between timed steps (outside the per-step event pairs); working set < L2",
                   "agents_without_flow_direction": frac_no_dir},
        "flow_fields_per_sec": (nf + nl) * world / ((groups["flow"] + groups["los"]) / args.steps * 1e-3) if (groups["flow"] + groups["los"]) > 0 else None,
        "fields_per_step": {"flow": nf * world, "los": nl * world},
        "phase_ms_per_step": {k: v / args.steps for k, v in groups.items()},
        "wall_s": t_wall,
        "gpu_launches": launches,
        "e2e": {"value": total_agents * args.steps / e2e_s, "unit": "agent-updates/s",
                "h2d_bytes_per_step": int(n_total * capi.AGENT.itemsize + len(W["flocks"]) * capi.FLOCK.itemsize + nwork * 4),
                "d2h_bytes_per_step": int(nwork * 8)},
        "roofline": {"bound": "hbm", "kernel": {"velocity": "k_agent_velocity", "cohesion": "k_cohesion", "flow": "k_flow_unit",
                                                  "los": "k_los_b", "index": "k_cell_*", "vdes": "k_desired_velocity"}[dom],
                     "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm, "traffic": traffic,
                     "peak_source": peak_src + " (MEASURED_PEAKS.json hbm_gbs)" if peak_src == "measured" else "fallback 6650 GB/s",
                     "algorithmic_bytes_per_launch": per_launch[dom], "ms_per_launch": dom_ms,
                     "note": "latency/ALU-bound by construction: 360 B of compulsory traffic per agent update (SURVEY.md 8d)"},
        "clocks": sampler.summary() if rank == 0 else None,
    }
    if rank == 0:
        k10, k30 = neighbour_stats(W["agents"])
        result["config"]["k10_mean"] = k10; result["config"]["k30_mean"] = k30
        if world == 1 and not args.no_cpu_baseline:
            result["cpu_baseline"] = cpu_baseline(pf, budget_s=args.cpu_budget)
        print(json.dumps(result), flush=True)
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


def cpu_sample(pf, budget_s):
    """Bounded CPU sample of the same workload: one flock pair of the C2 population (6250 + 6250 agents,
    radii 1.5 / 3.0, same spawn-disc generator) on a 4x4-chunk map of the same generator, plus the
    destination-chunk flow + LOS fields of 1024 seeded goals. Returns (agent_updates_per_s, fields_per_s,
    cores, kind, description)."""
    synth, capi = pf.synth, pf.capi
    kind, mod = _ref_or_port()
    cores = os.cpu_count() or 1
    cw = 4
    p = synth.make_map(cw, cw, MAP_SEED)
    cost = synth.cost_from_pathable(p, cw, cw)
    a = synth.make_agents(cost, cw, cw, 12_500, 2, MAP_SEED, radius=np.array([1.5, 3.0], np.float32), spacing=2.2, hz=HZ)
    rng = np.random.default_rng(1)
    n_sample = 8192
    work = np.sort(rng.choice(12_500, n_sample, replace=False)).astype(np.uint32)
    tiles = synth.random_passable_tiles(cost, 1024, rng)
    reqs = np.array([[t[0] // cw, t[0] % cw, t[1], t[2]] for t in tiles], np.int32)
    t_start = time.perf_counter()
    if kind == "reference":
        ref = mod.RefMap(cw, cw, p)
        dest = [ref.dest_id((float(a["flock_target"][f][0]), float(a["flock_target"][f][1]))) for f in range(2)]
        ref.agents_set(a["pos"], a["prev_pos"], a["vel"], a["radius"], a["max_speed"], a["state"], a["flags"],
                       a["flock_of"], a["flock_target"], np.array(dest, np.uint32), hz=HZ)
        vdes = np.zeros((n_sample, 2), np.float32)
        d = a["flock_target"][a["flock_of"][work]] - a["pos"][work]
        vdes[:] = d / np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-6)
        ref.work_set(work, vdes, np.zeros(n_sample, np.uint8), a["speed"][work])
        # grow the sample until the budget is used
        _, secs = ref.velocity_work(cores)
        agents_per_s = n_sample / secs
        f_secs, _ = ref.fields_mt(0, reqs, cores)
        l_secs, _ = ref.fields_mt(1, reqs, cores)
        fields_per_s = 2 * len(reqs) / (f_secs + l_secs)
        ref.close()
        threads = cores
    else:
        om = mod.OracleMap(cw, cw, cost)
        aa = dict(a)
        d = a["flock_target"][a["flock_of"]] - a["pos"]
        aa["vdes"] = (d / np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-6)).astype(np.float32)
        rec, fl = capi.pack_agents(aa)
        w = mod.OracleWorld(om, rec, fl, HZ)
        t0 = time.perf_counter(); w.velocity_work(work); secs = time.perf_counter() - t0
        agents_per_s = n_sample / secs
        fr = np.concatenate([capi.tile_req((int(q[0]), int(q[1])), (int(q[2]), int(q[3]))) for q in reqs])
        lr = np.concatenate([capi.los_req((int(q[0]), int(q[1])), (int(q[0]), int(q[1]), int(q[2]), int(q[3]))) for q in reqs])
        t0 = time.perf_counter(); om.flow_fields_update(fr); om.los_fields_create(lr); fs = time.perf_counter() - t0
        fields_per_s = 2 * len(reqs) / fs
        w.close()
        threads = 1
    desc = ("%d-agent sample of one radius-1.5 + one radius-3.0 flock (6250 each, C2 generator) on a 4x4-chunk map, "
            "move_velocity_work with the reference's equal-range split over %d threads; %d destination-chunk flow + %d LOS fields; "
            "%.1f s of CPU wall" % (n_sample, threads, len(reqs), len(reqs), time.perf_counter() - t_start))
    return agents_per_s, fields_per_s, threads, kind, desc


def cpu_baseline(pf, budget_s=20.0):
    v, f, cores, kind, desc = cpu_sample(pf, budget_s)
    return {"value": v, "unit": "agent-updates/s", "cores": cores, "kind": kind, "sample": desc,
            "flow_fields_per_sec": f, "host_cpus": os.cpu_count()}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pf = importlib.import_module("permafrost-engine_b200")
    vals, fvals = [], []
    t0 = time.perf_counter()
    info = None
    for k in range(args.warmup + args.steps):
        v, f, cores, kind, desc = cpu_sample(pf, args.cpu_budget)
        info = (cores, kind, desc)
        if k >= args.warmup:
            vals.append(v); fvals.append(f)
        if time.perf_counter() - t0 > 240 and len(vals) >= 1:
            break
    value = float(np.median(vals))
    cores, kind, desc = info
    world = int(os.environ.get("WORLD_SIZE", "1"))
    out = {
        "impl": "reference", "metric": "agent_updates_per_sec", "value": value, "unit": "agent-updates/s",
        "n_gpus": args.gpus, "steps": len(vals), "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD + ": 1024x1024-tile map (16x16 chunks), %d agents/GPU in %d flocks/GPU, %d flow-field goals/GPU "
                               "(every chunk connected to each goal), hz=20" % (AGENTS_PER_GPU, GOALS_PER_GPU, GOALS_PER_GPU),
                   "note": "CPU arm: throughput of the reference's own code on a bounded sample of that workload; "
                           "it does not scale with --gpus (rank 0 only, world=%d)" % world},
        "flow_fields_per_sec": float(np.median(fvals)),
        "cpu_baseline": {"value": value, "unit": "agent-updates/s", "cores": cores, "kind": kind, "sample": desc,
                         "host_cpus": os.cpu_count()},
        "e2e": {"value": value, "unit": "agent-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=20.0)
    ap.add_argument("--workload", default="C2", choices=["C2", "C3"],
                    help="C2 (default, the headline config): 100k agents / 16 goals per GPU; C3: 1M agents of radius 1.0 in "
                         "64 flocks / 64 goals (BASELINE.json configs[2]), extra evidence only")
    args = ap.parse_args()
    set_workload(args.workload)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
