"""Diagnostic (GPU box): the marching-crowd trajectory of tests/test_gpu_scale.py, free-running reference versus a reference
re-seeded from the device state before every tick. Prints per-tick deviation statistics."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import cases  # noqa: E402
import pfref  # noqa: E402

capi = cases.capi


def run(hz, nticks, resync):
    cw = 3
    p, cost, a = cases.agent_case(cw, 1500, 3, 3131, 0.03, 2.6)
    a["state"][:] = 0
    n = len(a["radius"])
    ms = np.zeros(n, capi.MOVESTATE)
    ms["next_pos"][:, 0] = a["pos"][:, 0]; ms["next_pos"][:, 2] = a["pos"][:, 1]
    ms["step"] = 1.0 / (20 // hz)
    ms["next_rot"] = cases.dir_quat(a["vel"] + 1e-6)
    ms["combat_facing"] = ms["next_rot"]
    ms["vel_hist"] = np.repeat(a["vel"][:, None, :], 14, axis=1)
    ref = pfref.RefMap(cw, cw, p)
    nav = capi.Nav(0)
    dest_ids = np.array([ref.dest_id((float(t[0]), float(t[1]))) for t in a["flock_target"]], np.uint32)
    ref.agents_set(a["pos"], a["prev_pos"], a["vel"], a["radius"], a["max_speed"], a["state"], a["flags"],
                   a["flock_of"], a["flock_target"], dest_ids, hz=hz)
    ref.movestate_set(ms["next_pos"][:, [0, 2]], ms["next_rot"], ms["step"], ms["left"], ms["vel_hist"], ms["vel_hist_idx"],
                      np.zeros(n, np.int32), np.zeros(n, np.int32), ms["combat_facing"])
    nav.map_create(cw, cw, 1); nav.map_upload_layer(0, cost); nav.map_build_nav(0); nav.route_build(0)
    nflocks = len(a["flock_target"])
    nav.pool_create(nflocks, nflocks * cw * cw)
    aa = dict(a); aa["flock_dest_index"] = np.arange(nflocks, dtype=np.int32)
    rec, fl = capi.pack_agents(aa)
    nav.agents_upload(rec, fl, hz)
    nav.agents_upload_movestate(ms)
    state = a["state"].copy()
    work = np.nonzero((state != 2) & (state != 4))[0].astype(np.uint32)
    ref.work_set(work, np.zeros((len(work), 2), np.float32), np.zeros(len(work), np.uint8), a["speed"][work])
    ref.desired_from_cache()
    got, gms = nav.agents_read_state(n)
    for tick in range(nticks):
        work = np.nonzero((state != 2) & (state != 4))[0].astype(np.uint32)
        est0 = ref.state_get(n)
        dpos = np.abs(got["pos"] - est0["pos"]).max(axis=1)
        if resync and tick:
            ref.agents_set(got["pos"], got["prev_pos"], got["velocity"], a["radius"], a["max_speed"], got["state"], a["flags"],
                           a["flock_of"], a["flock_target"], dest_ids, hz=hz)
            ref.movestate_set(gms["next_pos"][:, [0, 2]], gms["next_rot"], gms["step"], gms["left"], gms["vel_hist"], gms["vel_hist_idx"],
                              np.zeros(n, np.int32), np.zeros(n, np.int32), gms["combat_facing"])
        ref.work_set(work, np.zeros((len(work), 2), np.float32), np.zeros(len(work), np.uint8), a["speed"][work])
        evdes, elos = ref.desired_from_cache()
        evel, _ = ref.velocity_work(os.cpu_count())
        ref.update_and_apply()
        est = ref.state_get(n)
        nav.agents_set_work(work)
        for _ in range(8):
            nav.agents_tick(capi.TICK_VDES_FROM_POOL)
            nreq, nrep = nav.pool_repair()
            if nreq + nrep == 0:
                break
        nav.agents_tick(capi.TICK_VDES_FROM_POOL)
        vel = nav.agents_read_velocities(len(work))
        vpref, vdes, los = nav.agents_read_debug(len(work))
        re = cases.relerr(vel, evel)
        worst = int(np.argmax(re))
        print("hz %d resync %d tick %d: nwork %d, pos dev before %.2e, los mismatches %d, vdes dev %.2e, vel relerr max %.2e (>1e-5: %d, >1e-4: %d)"
              % (hz, resync, tick, len(work), dpos.max(), int((los != elos).sum()), np.abs(vdes - evdes).max(), re.max(),
                 int((re > 1e-5).sum()), int((re > 1e-4).sum())))
        if re.max() > 1e-5:
            u = work[worst]
            print("   worst uid %d: vel %s ref %s vdes %s ref %s pos dev %.2e" % (u, vel[worst], evel[worst], vdes[worst], evdes[worst], dpos[u]))
        nav.agents_compute_updates()
        nav.agents_apply_updates()
        nav.agents_rebuild_index()
        got, gms = nav.agents_read_state(n)
        state = est["state"]
        if not resync and (got["state"] != est["state"]).any():
            print("   state mismatch", np.nonzero(got["state"] != est["state"])[0][:10])
        if resync:
            state = got["state"].astype(state.dtype)
            e_pos = np.abs(got["pos"] - est["pos"]).max()
            print("   after apply: pos dev %.2e vel dev %.2e" % (e_pos, cases.relerr(got["velocity"], est["vel"]).max()))
    nav.close(); ref.close()


if __name__ == "__main__":
    for hz in (20, 10):
        run(hz, 12, False)
        run(hz, 12, True)
