"""ctypes binding of oracle/_ref/libpfref.so -- the UNMODIFIED reference compiled by
oracle/Makefile (`make ref`).  TEST INFRASTRUCTURE: imported only by tests/, the golden-vector
generator, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libpfref.so")

_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(LIB_PATH)
        L = _lib
        L.pfref_map_new.restype = C.c_void_p
        L.pfref_map_new.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_float, C.c_float]
        L.pfref_map_new_tiles.restype = C.c_void_p
        L.pfref_map_new_tiles.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_float, C.c_float]
        L.pfref_map_free.argtypes = [C.c_void_p]
        L.pfref_get_field.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.pfref_get_portals.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.pfref_get_portal_edges.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.pfref_flow_field_tile.argtypes = [C.c_void_p] + [C.c_int] * 7 + [C.c_void_p]
        L.pfref_flow_field_portal.argtypes = [C.c_void_p] + [C.c_int] * 8 + [C.c_void_p]
        L.pfref_flow_nearest_pathable.argtypes = [C.c_void_p] + [C.c_int] * 5 + [C.c_void_p]
        L.pfref_flow_island_to_nearest.argtypes = [C.c_void_p] + [C.c_int] * 9 + [C.c_void_p]
        L.pfref_cell_arrival_field.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                               C.c_int, C.c_void_p, C.c_void_p]
        L.pfref_group_arrival_field.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                                C.c_void_p, C.c_int, C.c_void_p]
        L.pfref_flow_field_zone.argtypes = [C.c_void_p] + [C.c_int] * 6 + [C.c_void_p]
        L.pfref_group_arrival_velocity.restype = C.c_int
        L.pfref_group_arrival_velocity.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.pfref_agents_set_factions.argtypes = [C.c_int, C.c_void_p]
        L.pfref_flow_field_entity.argtypes = [C.c_void_p] + [C.c_int] * 5 + [C.c_void_p]
        L.pfref_set_war.argtypes = [C.c_int, C.c_int, C.c_int]
        L.pfref_los_field_faction.argtypes = [C.c_void_p] + [C.c_int] * 8 + [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.pfref_los_field.argtypes = [C.c_void_p] + [C.c_int] * 7 + [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.pfref_request_path.argtypes = [C.c_void_p, C.c_int] + [C.c_float] * 4 + [C.c_void_p]
        L.pfref_request_path_attacking.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_float] * 4 + [C.c_void_p]
        L.pfref_dest_id.restype = C.c_uint32
        L.pfref_dest_id.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float]
        L.pfref_fc_get_flow.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.pfref_fc_get_los.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_void_p]
        L.pfref_fc_clear.argtypes = [C.c_void_p]
        L.pfref_desired_velocity.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p,
                                             C.c_float, C.c_float, C.c_void_p, C.c_void_p]
        L.pfref_blockers.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int, C.c_uint32]
        L.pfref_blockers_obb.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_uint32]
        L.pfref_update.argtypes = [C.c_void_p]
        L.pfref_clearpath.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        L.pfref_agents_set.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 8 + [C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.pfref_ents_in_circle.argtypes = [C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_int]
        L.pfref_work_set.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.pfref_velocity_work.argtypes = [C.c_int, C.c_int]
        L.pfref_velocity_work_mt.restype = C.c_double
        L.pfref_velocity_work_mt.argtypes = [C.c_int]
        L.pfref_work_get.argtypes = [C.c_int, C.c_void_p]
        L.pfref_movestate_set.argtypes = [C.c_int] + [C.c_void_p] * 9
        L.pfref_compute_updates.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.pfref_apply_velocity_patch.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.pfref_vpref.argtypes = [C.c_int, C.c_void_p]
        L.pfref_work_set_formation.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        L.pfref_movestate_ext_set.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        L.pfref_compute_updates_ext.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.pfref_desired_from_cache.argtypes = [C.c_void_p, C.c_void_p]
        L.pfref_update_and_apply.restype = C.c_int
        L.pfref_update_and_apply.argtypes = [C.c_void_p]
        L.pfref_state_get.argtypes = [C.c_int] + [C.c_void_p] * 5
        L.pfref_fields_mt.restype = C.c_double
        L.pfref_fields_mt.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class RefMap:
    """One reference nav context (N_NewCtxForMapData) over a synthetic pathable-tile grid."""

    def __init__(self, chunk_w, chunk_h, pathable=None, map_x=0.0, map_z=0.0, tiles=None):
        """pathable: u8[H32][W32] (all tiles FLAT, height 0)  -- or --
        tiles: int32[H32][W32][4] = {pathable, type, base_height, ramp_height}."""
        self.cw, self.ch = chunk_w, chunk_h
        self.map_x, self.map_z = float(map_x), float(map_z)
        if tiles is not None:
            tiles = np.ascontiguousarray(tiles, dtype=np.int32)
            assert tiles.shape == (chunk_h * 32, chunk_w * 32, 4)
            self.h = lib().pfref_map_new_tiles(chunk_w, chunk_h, _p(tiles), map_x, map_z)
        else:
            pathable = np.ascontiguousarray(pathable, dtype=np.uint8)
            assert pathable.shape == (chunk_h * 32, chunk_w * 32)
            self.h = lib().pfref_map_new(chunk_w, chunk_h, _p(pathable), map_x, map_z)
        if not self.h:
            raise RuntimeError("pfref_map_new failed")
        for a in range(15):                     # the harness's war matrix is process-global: every map starts at peace
            for b in range(a + 1, 15):
                lib().pfref_set_war(a, b, 0)

    def close(self):
        if self.h:
            lib().pfref_map_free(self.h)
            self.h = None

    def field(self, layer, kind):
        dt = np.uint8 if kind == 0 else np.uint16
        out = np.zeros((self.ch * self.cw, 64, 64), dtype=dt)
        lib().pfref_get_field(self.h, layer, kind, _p(out))
        return out

    def factions(self, layer=0):
        """u8[chunks][15][64][64] per-faction blocker refcounts (nav_data.h chunk->factions)"""
        out = np.zeros((self.ch * self.cw, 15, 64, 64), np.uint8)
        for f in range(15):
            tmp = np.zeros((self.ch * self.cw, 64, 64), np.uint8)
            lib().pfref_get_field(self.h, layer, 16 + f, _p(tmp))
            out[:, f] = tmp
        return out

    def cost_base(self, layer=0): return self.field(layer, 0)
    def blockers(self, layer=0): return self.field(layer, 1)
    def islands(self, layer=0): return self.field(layer, 2)
    def local_islands(self, layer=0): return self.field(layer, 3)

    def portals(self, layer=0):
        n = lib().pfref_get_portals(self.h, layer, None, 0)
        out = np.zeros((n, 10), dtype=np.int32)
        lib().pfref_get_portals(self.h, layer, _p(out), n)
        return out

    def portal_edges(self, layer, chunk_idx, portal_idx):
        out = np.zeros((64, 3), dtype=np.uint32)
        n = lib().pfref_get_portal_edges(self.h, layer, chunk_idx, portal_idx, _p(out), 64)
        return out[:n]

    def flow_tile(self, chunk, tile, layer=0, faction=0xF, inout=None):
        init = inout is None
        buf = np.zeros(4096, dtype=np.uint8) if init else np.ascontiguousarray(inout, dtype=np.uint8).reshape(-1).copy()
        lib().pfref_flow_field_tile(self.h, layer, chunk[0], chunk[1], tile[0], tile[1], faction, int(init), _p(buf))
        return buf.reshape(64, 64)

    def flow_nearest_pathable(self, chunk, start, inout, layer=0):
        buf = np.ascontiguousarray(inout, dtype=np.uint8).reshape(-1).copy()
        lib().pfref_flow_nearest_pathable(self.h, layer, chunk[0], chunk[1], start[0], start[1], _p(buf))
        return buf.reshape(64, 64)

    def flow_island_to_nearest(self, chunk, local_iid, inout, tile=None, portal=None, layer=0):
        """portal = (portal_idx, port_iid, next_iid) or tile = (r, c): the target that built the field"""
        buf = np.ascontiguousarray(inout, dtype=np.uint8).reshape(-1).copy()
        t = tile if tile is not None else (0, 0)
        pi = portal if portal is not None else (-1, 0xFFFF, 0xFFFF)
        lib().pfref_flow_island_to_nearest(self.h, layer, chunk[0], chunk[1], t[0], t[1], pi[0], pi[1], pi[2],
                                           int(local_iid), _p(buf))
        return buf.reshape(64, 64)

    def flow_portal(self, chunk, portal_idx, port_iid, next_iid, layer=0, faction=0xF, inout=None):
        init = inout is None
        buf = np.zeros(4096, dtype=np.uint8) if init else np.ascontiguousarray(inout, dtype=np.uint8).reshape(-1).copy()
        lib().pfref_flow_field_portal(self.h, layer, chunk[0], chunk[1], portal_idx, port_iid, next_iid,
                                      faction, int(init), _p(buf))
        return buf.reshape(64, 64)

    def cell_arrival_field(self, dim, target, center, enemies=0, overlay=None, fixup_start=None, layer=0):
        """N_CellArrivalFieldCreate [+ ...UpdateToNearestPathable]; absolute (r, c) tile coordinates."""
        t = np.asarray(target, np.int32); c = np.asarray(center, np.int32)
        ov = np.ascontiguousarray(overlay if overlay is not None else np.zeros((0, 2)), np.int32).reshape(-1, 2)
        st = None if fixup_start is None else np.asarray(fixup_start, np.int32)
        out = np.zeros((dim, dim // 2), np.uint8)
        lib().pfref_cell_arrival_field(self.h, dim, layer, int(enemies), _p(t), _p(c), _p(ov), len(ov), _p(st), _p(out))
        return out

    def group_arrival_field(self, dim, targets_xz, center_xz, enemies=0, overlay=None, layer=0):
        t = np.ascontiguousarray(targets_xz, np.float32).reshape(-1, 2); c = np.asarray(center_xz, np.float32)
        ov = np.ascontiguousarray(overlay if overlay is not None else np.zeros((0, 2)), np.int32).reshape(-1, 2)
        out = np.zeros((dim, dim // 2), np.uint8)
        lib().pfref_group_arrival_field(self.h, dim, layer, int(enemies), _p(t), len(t), _p(c), _p(ov), len(ov), _p(out))
        return out

    def flow_field_zone(self, chunk, centre, radius, layer=0):
        """N_FlowFieldUpdate(TARGET_ZONE) for one chunk; centre = absolute (r, c)"""
        out = np.zeros((64, 64), np.uint8)
        lib().pfref_flow_field_zone(self.h, layer, chunk[0], chunk[1], int(centre[0]), int(centre[1]), int(radius), _p(out))
        return out

    def group_arrival_velocity(self, centre_xz, radius, pos_xz, layer=0):
        """request + await the zone fields in reach, then N_DesiredGroupArrivalVelocity per position
        -> (vel[n, 2], flags[n]: bit0 ok, bit1 at_slot, chunk fields built)"""
        pos = np.ascontiguousarray(pos_xz, np.float32).reshape(-1, 2); c = np.ascontiguousarray(centre_xz, np.float32)
        vel = np.zeros((len(pos), 2), np.float32); fl = np.zeros(len(pos), np.uint8)
        nb = lib().pfref_group_arrival_velocity(self.h, layer, _p(c), int(radius), len(pos), _p(pos), _p(vel), _p(fl))
        return vel, fl, nb

    def agents_set_factions(self, factions):
        f = np.ascontiguousarray(factions, np.int32)
        lib().pfref_agents_set_factions(len(f), _p(f))

    def flow_field_entity(self, chunk, uid, layer=0):
        """N_FlowFieldUpdate(TARGET_ENTITY) around uploaded agent `uid` (surround fields)"""
        out = np.zeros((64, 64), np.uint8)
        lib().pfref_flow_field_entity(self.h, layer, chunk[0], chunk[1], 0, int(uid), _p(out))
        return out

    def flow_field_enemies(self, chunk, faction, layer=0):
        """N_FlowFieldUpdate(TARGET_ENEMIES): towards the nearest enemy of `faction` among the uploaded agents"""
        out = np.zeros((64, 64), np.uint8)
        lib().pfref_flow_field_entity(self.h, layer, chunk[0], chunk[1], 1, int(faction), _p(out))
        return out

    def set_war(self, a, b, at_war=True):
        lib().pfref_set_war(a, b, int(at_war))

    def los(self, chunk, target_td, layer=0, prev=None, prev_chunk=(0, 0), faction=0xF):
        out = np.zeros(4096, dtype=np.uint8)
        pv = None if prev is None else np.ascontiguousarray(prev, dtype=np.uint8).reshape(-1)
        if faction != 0xF:
            lib().pfref_los_field_faction(self.h, faction, layer, chunk[0], chunk[1], target_td[0], target_td[1],
                                          target_td[2], target_td[3], _p(pv), prev_chunk[0], prev_chunk[1], _p(out))
            return out.reshape(64, 64)
        lib().pfref_los_field(self.h, layer, chunk[0], chunk[1], target_td[0], target_td[1], target_td[2],
                              target_td[3], _p(pv), prev_chunk[0], prev_chunk[1], _p(out))
        return out.reshape(64, 64)

    def request_path(self, src, dst, layer=0, faction=0xF):
        did = C.c_uint32(0)
        if faction != 0xF:
            ok = lib().pfref_request_path_attacking(self.h, layer, faction, src[0], src[1], dst[0], dst[1], C.byref(did))
        else:
            ok = lib().pfref_request_path(self.h, layer, src[0], src[1], dst[0], dst[1], C.byref(did))
        return bool(ok), did.value

    def dest_id(self, dst, layer=0):
        return lib().pfref_dest_id(self.h, layer, dst[0], dst[1])

    def fc_flow(self, dest_id, chunk):
        out = np.zeros(4096, dtype=np.uint8)
        ffid = C.c_uint64(0)
        ok = lib().pfref_fc_get_flow(self.h, dest_id, chunk[0], chunk[1], _p(out), C.byref(ffid))
        return (out.reshape(64, 64), ffid.value) if ok else (None, None)

    def fc_los(self, dest_id, chunk):
        out = np.zeros(4096, dtype=np.uint8)
        ok = lib().pfref_fc_get_los(self.h, dest_id, chunk[0], chunk[1], _p(out))
        return out.reshape(64, 64) if ok else None

    def fc_clear(self):
        lib().pfref_fc_clear(self.h)

    def desired_velocity(self, dest_id, pos, los_pos, dest_xz):
        pos = np.ascontiguousarray(pos, dtype=np.float32)
        los_pos = np.ascontiguousarray(los_pos, dtype=np.float32)
        n = pos.shape[0]
        vdes = np.zeros((n, 2), dtype=np.float32)
        los = np.zeros(n, dtype=np.uint8)
        lib().pfref_desired_velocity(self.h, dest_id, n, _p(pos), _p(los_pos), dest_xz[0], dest_xz[1],
                                     _p(vdes), _p(los))
        return vdes, los

    def blockers_incref(self, x, z, radius, faction=0, flags=0):
        lib().pfref_blockers(self.h, 1, x, z, radius, faction, flags)

    def blockers_decref(self, x, z, radius, faction=0, flags=0):
        lib().pfref_blockers(self.h, 0, x, z, radius, faction, flags)

    def blockers_obb(self, corners_xz, incref=True, faction=0, flags=0):
        c = np.ascontiguousarray(corners_xz, np.float32).reshape(8)
        lib().pfref_blockers_obb(self.h, int(incref), _p(c), faction, flags)

    def update(self):
        lib().pfref_update(self.h)

    def agents_set(self, pos, prev_pos, vel, radius, max_speed, state, flags, flock_of,
                   flock_target, flock_dest, hz=20):
        f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
        n = len(radius)
        self._keep = [f32(pos), f32(prev_pos), f32(vel), f32(radius), f32(max_speed),
                      np.ascontiguousarray(state, dtype=np.int32), np.ascontiguousarray(flags, dtype=np.uint32),
                      np.ascontiguousarray(flock_of, dtype=np.int32), f32(flock_target),
                      np.ascontiguousarray(flock_dest, dtype=np.uint32)]
        k = self._keep
        lib().pfref_agents_set(self.h, n, _p(k[0]), _p(k[1]), _p(k[2]), _p(k[3]), _p(k[4]), _p(k[5]), _p(k[6]),
                               _p(k[7]), len(k[9]), _p(k[8]), _p(k[9]), hz)

    def ents_in_circle(self, x, z, r, maxout=512):
        out = np.zeros(maxout, dtype=np.uint32)
        n = lib().pfref_ents_in_circle(x, z, r, _p(out), maxout)
        return out[:n].copy()

    def work_set(self, uids, vdes, has_los, speed):
        uids = np.ascontiguousarray(uids, dtype=np.uint32)
        vdes = np.ascontiguousarray(vdes, dtype=np.float32)
        has_los = np.ascontiguousarray(has_los, dtype=np.uint8)
        speed = np.ascontiguousarray(speed, dtype=np.float32)
        self._nwork = len(uids)
        lib().pfref_work_set(self._nwork, _p(uids), _p(vdes), _p(has_los), _p(speed))

    def velocity_work(self, nthreads=1):
        secs = lib().pfref_velocity_work_mt(nthreads)
        out = np.zeros((self._nwork, 2), dtype=np.float32)
        lib().pfref_work_get(self._nwork, _p(out))
        return out, secs

    def movestate_set(self, next_pos, next_rot, step, left, vel_hist, vel_hist_idx, wait_prev, wait_ticks,
                      combat_facing):
        """interpolation / orientation / wait fields of struct movestate (movement.c:150-215), by uid"""
        f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        k = [f32(next_pos), f32(next_rot), f32(step), i32(left), f32(vel_hist), i32(vel_hist_idx), i32(wait_prev),
             i32(wait_ticks), f32(combat_facing)]
        lib().pfref_movestate_set(len(k[2]), *[_p(a) for a in k])

    def compute_updates(self, new_vel):
        """entity_compute_update (movement.c:2303) per work item -> (ints[n,4], floats[n,28]);
        layout in oracle/ref_harness.c:pfref_compute_updates"""
        new_vel = np.ascontiguousarray(new_vel, dtype=np.float32)
        oi = np.zeros((self._nwork, 4), np.int32)
        of = np.zeros((self._nwork, 28), np.float32)
        lib().pfref_compute_updates(self._nwork, _p(new_vel), _p(oi), _p(of))
        return oi, of

    def apply_velocity_patch(self, next_velocity, flags):
        nv = np.ascontiguousarray(next_velocity, dtype=np.float32)
        fl = np.ascontiguousarray(flags, dtype=np.int32)
        hist = np.zeros((self._nwork, 14, 2), np.float32)
        idx = np.zeros(self._nwork, np.int32)
        lib().pfref_apply_velocity_patch(self._nwork, _p(nv), _p(fl), _p(hist), _p(idx))
        return hist, idx

    def work_set_formation(self, form14, flags):
        """formation inputs per work item (pfref_work_set_formation)"""
        f = np.ascontiguousarray(form14, np.float32).reshape(self._nwork, 14); fl = np.ascontiguousarray(flags, np.uint32)
        lib().pfref_work_set_formation(self._nwork, _p(f), _p(fl))

    def movestate_ext_set(self, ints4, floats11):
        i = np.ascontiguousarray(ints4, np.int32); f = np.ascontiguousarray(floats11, np.float32)
        lib().pfref_movestate_ext_set(len(i), _p(i), _p(f))

    def compute_updates_ext(self, new_vel):
        """-> (ints[n,4], floats[n,28], extra[n,9] = next_dest[2], next_target_prev[2], next_target_dir[4], next_attack)"""
        new_vel = np.ascontiguousarray(new_vel, dtype=np.float32)
        oi = np.zeros((self._nwork, 4), np.int32); of = np.zeros((self._nwork, 28), np.float32); ox = np.zeros((self._nwork, 9), np.float32)
        lib().pfref_compute_updates_ext(self._nwork, _p(new_vel), _p(oi), _p(of), _p(ox))
        return oi, of, ox

    def desired_from_cache(self):
        """compute_los_state + compute_desired_velocity (movement.c:4129, 4163) on the work list -> (vdes, los)"""
        vdes = np.zeros((self._nwork, 2), np.float32); los = np.zeros(self._nwork, np.uint8)
        lib().pfref_desired_from_cache(_p(vdes), _p(los))
        return vdes, los

    def update_and_apply(self):
        """entity_compute_update + entity_apply_update for the work list, then publish the next snapshot"""
        return lib().pfref_update_and_apply(self.h)

    def state_get(self, n):
        pos = np.zeros((n, 2), np.float32); prev = np.zeros((n, 2), np.float32); vel = np.zeros((n, 2), np.float32)
        st = np.zeros(n, np.int32); blk = np.zeros(n, np.int32)
        lib().pfref_state_get(n, _p(pos), _p(prev), _p(vel), _p(st), _p(blk))
        return dict(pos=pos, prev_pos=prev, vel=vel, state=st, blocking=blk)

    def vpref(self):
        out = np.zeros((self._nwork, 2), dtype=np.float32)
        lib().pfref_vpref(self._nwork, _p(out))
        return out

    def fields_mt(self, what, reqs, nthreads=1, layer=0, want_out=False):
        reqs = np.ascontiguousarray(reqs, dtype=np.int32)
        n = reqs.shape[0]
        out = np.zeros((n, 64, 64), dtype=np.uint8) if want_out else None
        secs = lib().pfref_fields_mt(self.h, layer, what, _p(reqs), n, nthreads, _p(out))
        return secs, out


def clearpath(self5, vpref, dyn, stat):
    self5 = np.ascontiguousarray(self5, dtype=np.float32)
    vpref = np.ascontiguousarray(vpref, dtype=np.float32)
    dyn = np.ascontiguousarray(dyn, dtype=np.float32).reshape(-1, 5)
    stat = np.ascontiguousarray(stat, dtype=np.float32).reshape(-1, 5)
    out = np.zeros(2, dtype=np.float32)
    lib().pfref_clearpath(_p(self5), _p(vpref), _p(dyn), dyn.shape[0], _p(stat), stat.shape[0], _p(out))
    return out
