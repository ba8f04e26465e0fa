// pfnav_agents.cu -- device field pool, device spatial index, and the per-agent velocity update.
//
// Replaces (reference file:line):
//   compute_los_state / compute_desired_velocity         src/game/movement.c:4129-4180
//     N_DesiredPointSeekVelocity, n_interpolated_flow_dir  src/navigation/nav.c:3468, 3407
//     N_HasDestLOS                                          src/navigation/nav.c:4026
//     M_Tile_DescForPoint2D / M_Tile_Bounds / RelativeDesc  src/map/tile.c:547, 356, 391
//   move_velocity_work                                    src/game/movement.c:3395-3466
//     point_seek_vpref :1870, point_seek_total_force :1745, arrive_force_point :1546,
//     cohesion_force :1653, separation_force :1690, nullify_impass_components :1831,
//     vec2_truncate :643, find_neighbours :2768
//   G_ClearPath_NewVelocity                               src/game/clearpath.c:694 (all of :123-418, 552-715)
//     C_InfiniteLineIntersection / C_RayRayIntersection2D   src/phys/collision.c:820, 854
//   G_Pos_EntsInCircleFrom -> bg_ent_inrange_circle       src/game/position.c:379, src/lib/public/bitmap_grid.h:1376
//
// Arithmetic mirrors the reference expression by expression (float storage, the same places
// promoted to double, IEEE sqrt/div, no FMA contraction: the TU is built with -fmad=false), so
// that every epsilon-threshold branch of ClearPath takes the same side as on the CPU.
//
// Kernels:
//   K5  k_cell_count / k_cell_scan / k_cell_scatter / k_cell_sort : counting sort of agents into
//       16-wu cells, in-cell order = descending uid (what insert-all + bg_cleanup leaves behind).
//   K6a k_desired_velocity : one thread per work item, bilinear flow blend + LOS bit from the pool.
//   K6b k_cohesion         : one thread per work item, O(|flock|) weighted centre of mass.
//   K6c k_agent_velocity   : ONE WARP PER AGENT. Lanes scan the hash cells, evaluate separation
//       terms, and split ClearPath's O(R^2) ray pairs / O(R) inside-PCR tests among themselves.
#include "pfnav_internal.cuh"
#include <algorithm>
#include <math.h>
#include <string.h>

#define EPS_F (1.0f / 1024)
#define FULL 0xffffffffu

// ------------------------------------------------------------------------------------------
// vec2 helpers (pf_math.c:58-94) -- float storage, sqrt evaluated in double then rounded, which
// is the correctly rounded float sqrt.
// ------------------------------------------------------------------------------------------
struct v2 { float x, z; };
__device__ __forceinline__ v2 v2_add(v2 a, v2 b) { return {a.x + b.x, a.z + b.z}; }
__device__ __forceinline__ v2 v2_sub(v2 a, v2 b) { return {a.x - b.x, a.z - b.z}; }
__device__ __forceinline__ v2 v2_scale(v2 a, float s) { return {a.x * s, a.z * s}; }
__device__ __forceinline__ float v2_dot(v2 a, v2 b) { return a.x * b.x + a.z * b.z; }
__device__ __forceinline__ float v2_len(v2 a) { return sqrtf(a.x * a.x + a.z * a.z); }
__device__ __forceinline__ v2 v2_normal(v2 a) { float l = v2_len(a); return {a.x / l, a.z / l}; }
// vec2_truncate (movement.c:643)
__device__ __forceinline__ v2 v2_truncate(v2 a, float max_len)
{
    if (v2_len(a) > max_len) { a = v2_normal(a); a = v2_scale(a, max_len); }
    return a;
}

// ------------------------------------------------------------------------------------------
// Map / pool views
// ------------------------------------------------------------------------------------------
struct MapView {
    const uint8_t *cost; const uint16_t *blk;   // [layer][H64][W64]
    int W64, H64, chunk_w, chunk_h;
    float map_x, map_z;
};
struct PoolView {
    const int32_t *slot;     // [ndests][chunks]
    const uint8_t *flow;     // [max][4096]
    const uint8_t *los;      // [max][4096] ; first byte of a never-written LOS slot region is valid 0s
    const uint8_t *has;      // [max] bit0 flow present, bit1 LOS present
    int ndests;
    uint32_t *touch;         // [max] last tick that read the slot (LRU stamp), may be NULL
    uint32_t tick_no;
};
struct GridView {
    const uint32_t *cell_start, *cell_count;
    const int32_t *ix, *iy; const uint32_t *id;
    int grid_w, grid_h; int32_t origin_x, origin_y;
};

struct tile_desc { int chunk_r, chunk_c, tile_r, tile_c; };

// M_Tile_DescForPoint2D with the nav resolution (tile.c:547; n_res nav.c:240): returns false
// outside the map box.
__device__ __forceinline__ bool desc_for_point(const MapView &m, float px, float pz, tile_desc &out)
{
    const float width = (float)(m.chunk_w * 256), height = (float)(m.chunk_h * 256);
    if (px > m.map_x || px < m.map_x - width) return false;
    if (pz < m.map_z || pz > m.map_z + height) return false;
    int chunk_r = (int)(fabs((double)(m.map_z - pz)) / 256.0);
    int chunk_c = (int)(fabs((double)(m.map_x - px)) / 256.0);
    chunk_r = min(max(chunk_r, 0), m.chunk_h - 1);
    chunk_c = min(max(chunk_c, 0), m.chunk_w - 1);
    const float base_x = m.map_x - (float)chunk_c * 256.0f;
    const float base_z = m.map_z + (float)chunk_r * 256.0f;
    int tile_r = (int)(fabs((double)(base_z - pz)) / 4);
    int tile_c = (int)(fabs((double)(base_x - px)) / 4);
    out.chunk_r = chunk_r; out.chunk_c = chunk_c;
    out.tile_r = min(max(tile_r, 0), 63);
    out.tile_c = min(max(tile_c, 0), 63);
    return true;
}

// M_NavPositionPathable / M_NavPositionBlocked (map.c:817, 831): both false outside the map box.
__device__ __forceinline__ void probe_tile(const MapView &m, int layer, float px, float pz, bool &pathable,
                                           bool &blocked)
{
    tile_desc td;
    pathable = false; blocked = false;
    if (!desc_for_point(m, px, pz, td)) return;
    const size_t off = ((size_t)layer * m.H64 + td.chunk_r * 64 + td.tile_r) * m.W64 + td.chunk_c * 64 + td.tile_c;
    pathable = m.cost[off] != 0xFF;
    blocked = m.blk[off] > 0;
}

// Entity_NavLayerWithRadius (entity.c:554)
__device__ __forceinline__ int nav_layer_for(uint32_t flags, float radius)
{
    const bool water = flags & PFNAV_FLAG_WATER, air = flags & PFNAV_FLAG_AIR;
    const int base = water ? 4 : air ? 8 : 0;
    if (radius >= 15.0f) return base + 3;
    if (radius >= 10.0f) return base + 2;
    if (radius >= 5.0f) return base + 1;
    return base;
}

// N_FlowDir (field.c:2429): 1.0f / sqrt(2.0f) is a double division rounded to float
__device__ __forceinline__ v2 flow_dir_vec(int dir)
{
    const float d = (float)(1.0 / 1.4142135623730951);
    switch (dir) {
    case 1: return {d, -d};
    case 2: return {0.0f, -1.0f};
    case 3: return {-d, -d};
    case 4: return {1.0f, 0.0f};
    case 5: return {-1.0f, 0.0f};
    case 6: return {d, d};
    case 7: return {0.0f, 1.0f};
    case 8: return {-d, d};
    default: return {0.0f, 0.0f};
    }
}

// ------------------------------------------------------------------------------------------
// K6a: desired velocity + LOS from the device field pool
// ------------------------------------------------------------------------------------------
// ent_desired_velocity (movement.c:1466) decides per movement state where the desired velocity comes from:
//   TURNING                      zero
//   SEEK_ENEMIES                 N_DesiredEnemySeekVelocity (nav.c:3603): direction of the own tile in the TARGET_ENEMIES
//                                field of its chunk -- pool destination aux_dest1 - 1 -- NOT interpolated
//   SURROUND_ENTITY              N_DesiredSurroundVelocity (nav.c:3687) out of the TARGET_ENTITY field the same way while
//                                ms->using_surround_field (and the target exists); the point-seek field otherwise
//   ARRIVING_TO_CELL             the caller's cell_arrival_vdes (formation.c owns the cell fields)
//   everything else              N_DesiredPointSeekVelocity (nav.c:3468)
// and compute_los_state (movement.c:4129): N_HasDestLOS at prev_pos for entities in a flock, except surround-field users.
__global__ void k_desired_velocity(MapView m, PoolView pool, const pfnav_agent *__restrict__ agents,
                                   const pfnav_flock *__restrict__ flocks, const uint32_t *__restrict__ work,
                                   int nwork, float2 *__restrict__ vdes_out, uint8_t *__restrict__ los_out,
                                   uint32_t *__restrict__ miss_count, const pfnav_movestate_ext *__restrict__ ext,
                                   const pfnav_formation_in *__restrict__ form)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= nwork) return;
    const uint32_t uid = work[w];
    const pfnav_agent a = agents[uid];
    v2 vdes = {0.0f, 0.0f};
    uint8_t los = 0;
    int dest = (a.flock >= 0) ? flocks[a.flock].dest : -1;
    bool field_dir_only = false, want_los = a.flock >= 0, want_vdes = true;
    if (a.state == PFNAV_STATE_TURNING) want_vdes = false;
    else if (a.state == PFNAV_STATE_ARRIVING_TO_CELL) {
        want_vdes = false;
        if (form) vdes = {form[uid].cell_arrival_vdes[0], form[uid].cell_arrival_vdes[1]};
    } else if (a.state == PFNAV_STATE_SEEK_ENEMIES) {
        dest = (int)a.aux_dest1 - 1; field_dir_only = true; want_los = false;
    } else if (a.state == PFNAV_STATE_SURROUND_ENTITY && ext && ext[uid].surround_target_uid != 0xffffffffu &&
               ext[uid].using_surround_field) {
        dest = (int)a.aux_dest1 - 1; field_dir_only = true; want_los = false;
    }
    if (field_dir_only) {
        tile_desc t;
        if (dest >= 0 && dest < pool.ndests && desc_for_point(m, a.pos[0], a.pos[1], t)) {
            const int s = pool.slot[(size_t)dest * m.chunk_w * m.chunk_h + t.chunk_r * m.chunk_w + t.chunk_c];
            if (s < 0 || !(pool.has[s] & 1)) atomicAdd(miss_count, 1u);
            else {
                vdes = flow_dir_vec(pool.flow[(size_t)s * 4096 + t.tile_r * 64 + t.tile_c] & 0xF);
                if (pool.touch && pool.touch[s] != pool.tick_no) pool.touch[s] = pool.tick_no;
            }
        }
        vdes_out[w] = make_float2(vdes.x, vdes.z);
        los_out[w] = 0;
        return;
    }
    const int pdest = (a.flock >= 0) ? flocks[a.flock].dest : -1;
    if (want_los && pdest >= 0 && pdest < pool.ndests) {
        tile_desc t;
        if (desc_for_point(m, a.prev_pos[0], a.prev_pos[1], t)) {
            const int s = pool.slot[(size_t)pdest * m.chunk_w * m.chunk_h + t.chunk_r * m.chunk_w + t.chunk_c];
            if (s >= 0 && (pool.has[s] & 2)) {
                los = pool.los[(size_t)s * 4096 + t.tile_r * 64 + t.tile_c] & 1;
                if (pool.touch && pool.touch[s] != pool.tick_no) pool.touch[s] = pool.tick_no;
            }
        }
    }
    if (!want_vdes) {
        vdes_out[w] = make_float2(vdes.x, vdes.z);
        los_out[w] = los;
        return;
    }
    if (dest >= 0 && dest < pool.ndests) {
        const int chunks = m.chunk_w * m.chunk_h;
        const int32_t *slots = pool.slot + (size_t)dest * chunks;
        tile_desc t;
        // (N_HasDestLOS, sampled at prev_pos, was evaluated above: movement.c:4137)
        // ---- N_DesiredPointSeekVelocity at pos ----
        if (desc_for_point(m, a.pos[0], a.pos[1], t)) {
            const int s = slots[t.chunk_r * m.chunk_w + t.chunk_c];
            if (s < 0 || !(pool.has[s] & 1)) {
                atomicAdd(miss_count, 1u);      // the host planner must request this (dest, chunk)
            } else {
                const uint8_t *base_ff = pool.flow + (size_t)s * 4096;
                const int base_dir = base_ff[t.tile_r * 64 + t.tile_c] & 0xF;
                if (pool.touch && pool.touch[s] != pool.tick_no) pool.touch[s] = pool.tick_no;
                // n_interpolated_flow_dir (nav.c:3407)
                const float bx = (m.map_x - (float)(t.chunk_c * 256)) - (float)(t.tile_c * 4);
                const float bz = (m.map_z + (float)(t.chunk_r * 256)) + (float)(t.tile_r * 4);
                const float cx = bx - 4.0f / 2.0f, cz = bz + 4.0f / 2.0f;
                const float dx = a.pos[0] - cx, dz = a.pos[1] - cz;
                const int dc = (dx < 0.0f) ? 1 : -1;
                const int dr = (dz > 0.0f) ? 1 : -1;
                const float wc = (float)fmin(fabs((double)dx) / 4.0, 1.0);
                const float wr = (float)fmin(fabs((double)dz) / 4.0, 1.0);
                const int sdc[4] = {0, dc, 0, dc}, sdr[4] = {0, 0, dr, dr};
                const float sw[4] = {(1.0f - wc) * (1.0f - wr), wc * (1.0f - wr), (1.0f - wc) * wr, wc * wr};
                v2 acc = {0.0f, 0.0f};
                float wsum = 0.0f;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    if (sw[i] <= 0.0f) continue;
                    const int abs_r = t.chunk_r * 64 + t.tile_r + sdr[i], abs_c = t.chunk_c * 64 + t.tile_c + sdc[i];
                    if (abs_r < 0 || abs_r >= m.H64 || abs_c < 0 || abs_c >= m.W64) continue;
                    const int cr2 = abs_r >> 6, cc2 = abs_c >> 6;
                    const uint8_t *ff = base_ff;
                    if (cr2 != t.chunk_r || cc2 != t.chunk_c) {
                        const int s2 = slots[cr2 * m.chunk_w + cc2];
                        if (s2 < 0 || !(pool.has[s2] & 1)) continue;
                        ff = pool.flow + (size_t)s2 * 4096;
                    }
                    const int dir = ff[(abs_r & 63) * 64 + (abs_c & 63)] & 0xF;
                    if (dir == 0) continue;
                    const v2 fd = flow_dir_vec(dir);
                    acc = v2_add(acc, v2_scale(fd, sw[i]));
                    wsum += sw[i];
                }
                if (wsum < 1e-6f || v2_len(acc) < 1e-6f) vdes = flow_dir_vec(base_dir);
                else vdes = v2_normal(acc);
            }
        }
    }
    vdes_out[w] = make_float2(vdes.x, vdes.z);
    los_out[w] = los;
}

// ------------------------------------------------------------------------------------------
// K6b: cohesion_force (movement.c:1653). One thread per work item; members in ascending uid.
// ------------------------------------------------------------------------------------------
// member positions of every flock, contiguous in member-list order (one coalesced stream for k_cohesion)
__global__ void k_gather_flock_pos(const pf_record *__restrict__ rec, const uint32_t *__restrict__ flock_members, int n,
                                   float2 *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const pf_record r = rec[flock_members[i]];
    out[i] = make_float2(r.px, r.pz);
}

// scaled integer coordinates + cell of the position index (bitmap_grid.h: 16-wu cells)
__device__ __forceinline__ int32_t bg_scale(float x) { return __float2int_rn(x * 256.0f); }   // BG_SCALE_F
__device__ __forceinline__ int cell_of(int32_t i, int32_t origin, int n)
{
    int c = (i - origin) >> 12;
    return min(max(c, 0), n - 1);
}

// cohesion weight exp(-6 (|d| - 37.5) / 50) = 2^(|d| * (-0.12 log2 e) + 4.5 log2 e) on the SFU (rsqrt + ex2, flush-to-zero
// forms: no denormal fix-up code; a flushed weight is < 1e-38 of a neighbour's)
__device__ __forceinline__ float coh_weight(float dx, float dz)
{
    const float len2 = fmaxf(fmaf(dx, dx, dz * dz), 1e-30f);
    float r, w;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(len2));
    const float len = len2 * r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(w) : "f"(fmaf(len, -0.12f * 1.4426950408889634f, 4.5f * 1.4426950408889634f)));
    return w;
}

// the whole member list of the flock (the definition; used for small flocks and as the fall-back of the windowed pass)
__global__ void k_cohesion(const pf_record *__restrict__ rec, const int32_t *__restrict__ flock_of,
                           const uint32_t *__restrict__ flock_start, const float2 *__restrict__ member_pos,
                           const uint32_t *__restrict__ uids, const uint32_t *__restrict__ nuids_dev, int nuids_host,
                           float scaled_max_force, float2 *__restrict__ out_by_uid)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    const int nu = nuids_dev ? (int)*nuids_dev : nuids_host;
    if (w >= nu) return;
    const uint32_t uid = uids[w];
    const int fl = flock_of[uid];
    v2 ret = {0.0f, 0.0f};
    if (fl >= 0) {
        const pf_record self = rec[uid];
        const uint32_t b = flock_start[fl], e = flock_start[fl + 1];
        v2 com = {0.0f, 0.0f};
        // The reference's sum runs over the flock in khash bucket order (movement.c:1660), i.e. it is not
        // reproducible to the last bit anyway; what must hold is the 1e-4 velocity budget. The weight
        // exp(-6 (|d| - 37.5) / 50) is therefore evaluated with the SFU (rsqrt + ex2, relative error < 1e-6) and
        // the member itself is removed afterwards: its weight is exp(4.5) exactly as evaluated here for |d| = 0.
        const float self_scale = coh_weight(0.0f, 0.0f);
#pragma unroll 4
        for (uint32_t k = b; k < e; k++) {
            const float2 p = member_pos[k];
            const float scale = coh_weight(p.x - self.px, p.y - self.pz);
            com.x = fmaf(p.x, scale, com.x);
            com.z = fmaf(p.y, scale, com.z);
        }
        const uint32_t cnt = e - b - 1;
        if (cnt > 0) {
            com.x -= self.px * self_scale;
            com.z -= self.pz * self_scale;
            com = v2_scale(com, 1.0f / (float)cnt);
            ret = v2_sub(com, v2{self.px, self.pz});
            ret = v2_truncate(ret, scaled_max_force);
        }
    }
    out_by_uid[uid] = make_float2(ret.x, ret.z);
}

// cohesion_force (movement.c:1653) for big flocks. The weight of a member falls off as exp(-0.12 d): beyond COH_R = 230 wu
// a member weighs < 1.1e-12 of one standing next to the entity, so the sum over the members within COH_R equals the full
// sum to float precision WHENEVER the entity has company nearby. The pass walks the position index instead of the member
// list: 128 consecutive entries of the index (spatial neighbours: the index is cell-major) share one window of cell rows,
// the window's entries stream through shared memory in tiles, every thread accumulates its own entity's sum. An entity
// whose window holds too little weight to make the cut-off harmless (sum < N * exp(4.5 - 0.12 * 230) * 1e7, i.e. the
// dropped members could matter at the 1e-7 level) is handed to the full-list kernel instead, so the result never
// depends on the cut-off by more than float rounding. Index coordinates are the 1/256-wu integers of the position index.
#define COH_R 230.0f
#define COH_THREADS 128
__device__ __forceinline__ void coh_finish(float ax, float az, float wsum, float self_w, float six_f, float siy_f, float sx, float sz,
                                           uint32_t N, float thresh_unit, float scaled_max_force, uint32_t uid,
                                           float2 *__restrict__ out_by_uid, uint32_t *__restrict__ fallback,
                                           uint32_t *__restrict__ nfallback)
{
    if (N <= 1) out_by_uid[uid] = make_float2(0.0f, 0.0f);
    else if (wsum - self_w >= (float)N * thresh_unit) {
        // the entity's own index entry was summed too: take it out again (weight exp(4.5) at distance 0)
        v2 com = {ax - six_f * self_w, az - siy_f * self_w};
        com = v2_scale(com, 1.0f / (float)(N - 1));
        v2 ret = v2_sub(com, v2{sx, sz});
        ret = v2_truncate(ret, scaled_max_force);
        out_by_uid[uid] = make_float2(ret.x, ret.z);
    } else fallback[atomicAdd(nfallback, 1u)] = uid;
}
__global__ void __launch_bounds__(COH_THREADS)
k_cohesion_window(GridView g, const pf_record *__restrict__ rec, const int32_t *__restrict__ sfl,
                  const uint32_t *__restrict__ flock_start, int n, uint32_t lo, uint32_t hi, float scaled_max_force,
                  float2 *__restrict__ out_by_uid, uint32_t *__restrict__ fallback, uint32_t *__restrict__ nfallback,
                  float4 *__restrict__ part, int nsplit)
{
    __shared__ float4 tile[COH_THREADS];       // {x, z, flock id bits, -}
    __shared__ int bb[4];
    const int tid = threadIdx.x;
    const float thresh_unit = __expf(4.5f - 0.12f * COH_R) * 1e7f;
    for (int base = blockIdx.x * COH_THREADS; base < n; base += gridDim.x * COH_THREADS) {
        const int k = base + tid;
        uint32_t uid = 0; int fl = -1; bool active = false;
        float sx = 0.0f, sz = 0.0f, six_f = 0.0f, siy_f = 0.0f, self_w = 0.0f;
        int mycx = 0, mycy = 0;
        if (k < n) {
            uid = g.id[k]; fl = sfl[k];
            active = uid >= lo && uid < hi && fl >= 0;
            mycx = cell_of(g.ix[k], g.origin_x, g.grid_w); mycy = cell_of(g.iy[k], g.origin_y, g.grid_h);
            six_f = (float)g.ix[k] * (1.0f / 256.0f); siy_f = (float)g.iy[k] * (1.0f / 256.0f);
        }
        if (!__syncthreads_or(active)) continue;
        if (tid == 0) { bb[0] = 0x7fffffff; bb[1] = -1; bb[2] = 0x7fffffff; bb[3] = -1; }
        __syncthreads();
        if (active) {
            const pf_record self = rec[uid];
            sx = self.px; sz = self.pz;
            self_w = coh_weight(six_f - sx, siy_f - sz);       // what the loop below adds for the entity's own entry
            atomicMin(&bb[0], mycx); atomicMax(&bb[1], mycx); atomicMin(&bb[2], mycy); atomicMax(&bb[3], mycy);
        }
        __syncthreads();
        const int cy0 = bb[2], cy1 = bb[3];
        float ax = 0.0f, az = 0.0f, wsum = 0.0f;
        const int RC = (int)(COH_R / 16.0f) + 1;
        // One window per grid row of the block's entities: 128 consecutive index entries lie in one row, or wrap from the
        // end of one into the start of the next (then one window around both would span the whole map width).
        for (int cyv = cy0; cyv <= cy1; cyv++) {
            const bool mine = active && mycy == cyv;
            __syncthreads();
            if (tid == 0) { bb[0] = 0x7fffffff; bb[1] = -1; }
            __syncthreads();
            if (mine) { atomicMin(&bb[0], mycx); atomicMax(&bb[1], mycx); }
            __syncthreads();
            const int cx0 = bb[0], cx1 = bb[1];
            if (cx1 < 0) continue;                                  // nobody in this row (uniform: read from shared memory)
            for (int ry = max(cyv - RC, 0) + (int)blockIdx.y; ry <= min(cyv + RC, g.grid_h - 1); ry += nsplit) {
                const int gap = max(abs(ry - cyv) - 1, 0);
                const float dy = (float)gap * 16.0f;
                if (dy > COH_R) continue;
                const int nx = (int)(sqrtf(COH_R * COH_R - dy * dy) / 16.0f) + 1;
                const int x0 = max(cx0 - nx, 0), x1 = min(cx1 + nx, g.grid_w - 1);
                const uint32_t s0 = g.cell_start[ry * g.grid_w + x0], s1 = g.cell_start[ry * g.grid_w + x1 + 1];
                for (uint32_t t0 = s0; t0 < s1; t0 += COH_THREADS) {
                    const uint32_t e = t0 + tid;
                    if (e < s1)
                        tile[tid] = make_float4((float)g.ix[e] * (1.0f / 256.0f), (float)g.iy[e] * (1.0f / 256.0f), __int_as_float(sfl[e]), 0.0f);
                    __syncthreads();
                    const int nt = (int)min((uint32_t)COH_THREADS, s1 - t0);
                    if (mine) {
#pragma unroll 4
                        for (int j = 0; j < nt; j++) {
                            const float4 m = tile[j];
                            const float scale = __float_as_int(m.z) == fl ? coh_weight(m.x - sx, m.y - sz) : 0.0f;
                            ax = fmaf(m.x, scale, ax); az = fmaf(m.y, scale, az); wsum += scale;
                        }
                    }
                    __syncthreads();
                }
            }
        }
        if (nsplit > 1) {                    // a small population split over blockIdx.y: k_cohesion_finish adds the parts up
            if (k < n) part[(size_t)blockIdx.y * n + k] = make_float4(ax, az, wsum, 0.0f);
        } else if (active) {
            coh_finish(ax, az, wsum, self_w, six_f, siy_f, sx, sz, flock_start[fl + 1] - flock_start[fl], thresh_unit,
                       scaled_max_force, uid, out_by_uid, fallback, nfallback);
        }
    }
}

// the window rows of one 128-entry run were split over `nsplit` blocks (populations too small to fill the chip with one
// block per run): add the partial sums in split order and finish as the fused path does
__global__ void k_cohesion_finish(GridView g, const pf_record *__restrict__ rec, const int32_t *__restrict__ sfl,
                                  const uint32_t *__restrict__ flock_start, int n, uint32_t lo, uint32_t hi, float scaled_max_force,
                                  float2 *__restrict__ out_by_uid, uint32_t *__restrict__ fallback, uint32_t *__restrict__ nfallback,
                                  const float4 *__restrict__ part, int nsplit)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t uid = g.id[k];
    const int fl = sfl[k];
    if (!(uid >= lo && uid < hi && fl >= 0)) return;
    float ax = 0.0f, az = 0.0f, wsum = 0.0f;
    for (int sp = 0; sp < nsplit; sp++) { const float4 p = part[(size_t)sp * n + k]; ax += p.x; az += p.y; wsum += p.z; }
    const pf_record self = rec[uid];
    const float six_f = (float)g.ix[k] * (1.0f / 256.0f), siy_f = (float)g.iy[k] * (1.0f / 256.0f);
    const float self_w = coh_weight(six_f - self.px, siy_f - self.pz);
    coh_finish(ax, az, wsum, self_w, six_f, siy_f, self.px, self.pz, flock_start[fl + 1] - flock_start[fl],
               __expf(4.5f - 0.12f * COH_R) * 1e7f, scaled_max_force, uid, out_by_uid, fallback, nfallback);
}

// ------------------------------------------------------------------------------------------
// K5: spatial index build (bitmap_grid.h: 16-wu cells, scaled int32 coordinates)
// ------------------------------------------------------------------------------------------

__global__ void k_cell_count(const pf_record *__restrict__ rec, int n, GridView g, uint32_t *__restrict__ count)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int cx = cell_of(bg_scale(rec[i].px), g.origin_x, g.grid_w);
    const int cy = cell_of(bg_scale(rec[i].pz), g.origin_y, g.grid_h);
    atomicAdd(&count[cy * g.grid_w + cx], 1u);
}

// single-CTA exclusive scan (cells <= 1M)
__global__ void k_cell_scan(const uint32_t *__restrict__ count, uint32_t *__restrict__ start, int ncells)
{
    __shared__ uint32_t part[1024];
    const int tid = threadIdx.x, per = (ncells + 1023) / 1024;
    const int b = tid * per, e = min(b + per, ncells);
    uint32_t s = 0;
    for (int i = b; i < e; i++) s += count[i];
    part[tid] = s;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        uint32_t v = tid >= off ? part[tid - off] : 0;
        __syncthreads();
        part[tid] += v;
        __syncthreads();
    }
    uint32_t run = tid ? part[tid - 1] : 0;
    for (int i = b; i < e; i++) { start[i] = run; run += count[i]; }
    if (tid == 1023) start[ncells] = part[1023];
}

// the same scan for big grids, three launches: 4096 cells per block (4 per thread) -> block sums, their scan, apply
__device__ __forceinline__ uint32_t block_excl_scan_1024(uint32_t v, uint32_t *warp_tot, uint32_t &total)
{
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += o; }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        uint32_t w = warp_tot[lane], wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, wi, d); if (lane >= d) wi += o; }
        warp_tot[lane] = wi - w;                       // exclusive
        if (lane == 31) warp_tot[32] = wi;             // block total
    }
    __syncthreads();
    total = warp_tot[32];
    return warp_tot[wid] + incl - v;
}

__global__ void __launch_bounds__(1024) k_cell_scan_sum(const uint32_t *__restrict__ count, uint32_t *__restrict__ part, int ncells)
{
    __shared__ uint32_t wt[33];
    const int i0 = (blockIdx.x * 1024 + threadIdx.x) * 4;
    uint32_t v = 0;
    for (int j = 0; j < 4; j++) if (i0 + j < ncells) v += count[i0 + j];
    uint32_t total;
    block_excl_scan_1024(v, wt, total);
    if (threadIdx.x == 0) part[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) k_cell_scan_parts(uint32_t *__restrict__ part, int nb, uint32_t *__restrict__ start, int ncells)
{
    __shared__ uint32_t wt[33];
    const uint32_t v = (int)threadIdx.x < nb ? part[threadIdx.x] : 0u;
    uint32_t total;
    const uint32_t ex = block_excl_scan_1024(v, wt, total);
    if ((int)threadIdx.x < nb) part[threadIdx.x] = ex;
    if (threadIdx.x == 0) start[ncells] = total;
}

__global__ void __launch_bounds__(1024) k_cell_scan_apply(const uint32_t *__restrict__ count, const uint32_t *__restrict__ part,
                                                          uint32_t *__restrict__ start, int ncells)
{
    __shared__ uint32_t wt[33];
    const int i0 = (blockIdx.x * 1024 + threadIdx.x) * 4;
    uint32_t c[4], v = 0;
    for (int j = 0; j < 4; j++) { c[j] = i0 + j < ncells ? count[i0 + j] : 0u; v += c[j]; }
    uint32_t total;
    uint32_t run = part[blockIdx.x] + block_excl_scan_1024(v, wt, total);
    for (int j = 0; j < 4; j++) if (i0 + j < ncells) { start[i0 + j] = run; run += c[j]; }
}

__global__ void k_cell_scatter(const pf_record *__restrict__ rec, int n, GridView g,
                               const uint32_t *__restrict__ start, uint32_t *__restrict__ fill,
                               uint32_t *__restrict__ sid)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int cx = cell_of(bg_scale(rec[i].px), g.origin_x, g.grid_w);
    const int cy = cell_of(bg_scale(rec[i].pz), g.origin_y, g.grid_h);
    const int c = cy * g.grid_w + cx;
    const uint32_t slot = start[c] + atomicAdd(&fill[c], 1u);
    sid[slot] = (uint32_t)i;
}

// per-cell: order ids DESCENDING (LIFO overflow chain drained by bg_cleanup, bitmap_grid.h:1110,
// 1477), then materialise the scaled coordinates next to them.
__global__ void k_cell_sort(const pf_record *__restrict__ rec, GridView g, const uint32_t *__restrict__ start,
                            uint32_t *__restrict__ sid, int32_t *__restrict__ six, int32_t *__restrict__ siy,
                            int ncells, const int32_t *__restrict__ flock_of, int32_t *__restrict__ sfl)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncells) return;
    const uint32_t b = start[c], e = start[c + 1];
    for (uint32_t i = b + 1; i < e; i++) {      // insertion sort, descending
        const uint32_t v = sid[i];
        uint32_t j = i;
        while (j > b && sid[j - 1] < v) { sid[j] = sid[j - 1]; j--; }
        sid[j] = v;
    }
    for (uint32_t i = b; i < e; i++) {
        const pf_record r = rec[sid[i]];
        six[i] = bg_scale(r.px);
        siy[i] = bg_scale(r.pz);
        sfl[i] = flock_of[sid[i]];
    }
}

// bg_ent_inrange_circle (bitmap_grid.h:1376): visit hits in the reference's order and hand them,
// 32 candidates at a time, to `emit(hit, id)` (called convergently by all lanes; returns a
// warp-uniform "stop").
template <class F>
__device__ __forceinline__ void grid_query(const GridView &g, float x, float z, float range, uint32_t lane,
                                           F &&emit)
{
    const int32_t icx = bg_scale(x), icy = bg_scale(z), ir = bg_scale(range);
    const long long ir2 = (long long)ir * ir;
    const int32_t imnx = icx - ir, imxx = icx + ir, imny = icy - ir, imxy = icy + ir;
    // _bg_cell_extent (bitmap_grid.h:1236)
    if (imxx < g.origin_x || imxy < g.origin_y) return;
    if (imnx >= g.origin_x + (g.grid_w << 12) || imny >= g.origin_y + (g.grid_h << 12)) return;
    const int cx_lo = max((imnx - g.origin_x) >> 12, 0), cx_hi = min((imxx - g.origin_x) >> 12, g.grid_w - 1);
    const int cy_lo = max((imny - g.origin_y) >> 12, 0), cy_hi = min((imxy - g.origin_y) >> 12, g.grid_h - 1);
    // wide-query fast path: whole pool in pool order == cells in row-major order
    const bool wide = (long long)(cx_hi - cx_lo + 1) * (cy_hi - cy_lo + 1) * 4 >= (long long)g.grid_w * g.grid_h * 3;
    const int ax_lo = wide ? 0 : cx_lo, ax_hi = wide ? g.grid_w - 1 : cx_hi;
    const int ay_lo = wide ? 0 : cy_lo, ay_hi = wide ? g.grid_h - 1 : cy_hi;
    const int cstep = wide ? (1 << 30) : 8;                    // one "coarse block" spans everything when wide
    const bool small = !wide && ir >= 0 && ir <= 12000;         // <= 46 wu
    const int cyc_lo = wide ? 0 : ay_lo >> 3, cyc_hi = wide ? 0 : ay_hi >> 3;
    const int cxc_lo = wide ? 0 : ax_lo >> 3, cxc_hi = wide ? 0 : ax_hi >> 3;
    for (int cyc = cyc_lo; cyc <= cyc_hi; cyc++) {
        for (int cxc = cxc_lo; cxc <= cxc_hi; cxc++) {
            const int fy0 = wide ? ay_lo : max(cyc * cstep, ay_lo), fy1 = wide ? ay_hi + 1 : min(cyc * cstep + cstep, ay_hi + 1);
            const int fx0 = wide ? ax_lo : max(cxc * cstep, ax_lo), fx1 = wide ? ax_hi + 1 : min(cxc * cstep + cstep, ax_hi + 1);
            for (int fy = fy0; fy < fy1; fy++) {
                // the cells fx0..fx1-1 of one fine row are adjacent in the cell-major index: one contiguous run of entries,
                // visited in ascending cell x, in-cell order -- the reference's order
                const uint32_t b = g.cell_start[fy * g.grid_w + fx0], e = g.cell_start[fy * g.grid_w + fx1];
                for (uint32_t k0 = b; k0 < e; k0 += 32) {
                    const uint32_t k = k0 + lane;
                    bool hit = false;
                    uint32_t id = 0;
                    if (k < e) {
                        if (small) {        // |d| <= range + one cell on either side: the squares fit 32 bits
                            const int32_t dx = g.ix[k] - icx, dy = g.iy[k] - icy;
                            hit = abs(dx) <= ir && abs(dy) <= ir && dx * dx + dy * dy <= (int32_t)ir2;
                        } else {
                            const long long dx = (long long)g.ix[k] - icx, dy = (long long)g.iy[k] - icy;
                            hit = dx * dx + dy * dy <= ir2;
                        }
                        id = g.id[k];
                    }
                    if (emit(hit, id)) return;
                }
            }
        }
    }
}

__device__ __forceinline__ int filter_garrisoned(uint32_t *ids, int count, const pf_record *__restrict__ rec, uint32_t lane);

__global__ void k_ents_in_circle(GridView g, float x, float z, float range, uint32_t *out, int maxout, int *out_n,
                                 const pf_record *__restrict__ rec, int filter_garr)
{
    const uint32_t lane = threadIdx.x & 31;
    int written = 0;
    grid_query(g, x, z, range, lane, [&](bool hit, uint32_t id) -> bool {
        const uint32_t m = __ballot_sync(FULL, hit);
        const int rank = __popc(m & ((1u << lane) - 1));
        if (hit && written + rank < maxout) out[written + rank] = id;
        written += __popc(m);
        return written >= maxout;
    });
    written = min(written, maxout);
    __syncwarp();
    if (filter_garr) written = filter_garrisoned(out, written, rec, lane);      // position.c:379
    if (lane == 0) *out_n = written;
}

// filter_garrisoned (position.c:100-119): swap-remove from the back, applied by G_Pos_EntsInCircleFrom
// (position.c:379) AFTER the raw query was cut at maxout; it reorders the survivors. Serial on lane 0
// (only runs when the population holds garrisoned entities at all); returns the new count to all lanes.
__device__ __forceinline__ int filter_garrisoned(uint32_t *ids, int count, const pf_record *__restrict__ rec, uint32_t lane)
{
    int ret = count;
    if (lane == 0) {
        for (int i = count - 1; i >= 0; i--) {
            if (rec[ids[i]].state_flags & PFNAV_FLAG_GARRISONED) { ids[i] = ids[ret - 1]; ret--; }
        }
    }
    ret = __shfl_sync(FULL, ret, 0);
    __syncwarp();
    return ret;
}

// ------------------------------------------------------------------------------------------
// ClearPath pieces
// ------------------------------------------------------------------------------------------
struct ray { v2 point, dir; };
struct cp_ent { v2 pos, vel; float radius; };

#define QNAN __int_as_float(0x7fc00000)

// slope used by C_InfiniteLineIntersection (collision.c:822): NaN marks a (near-)vertical line
__device__ __forceinline__ float line_slope(v2 dir) { return fabsf(dir.x) < EPS_F ? QNAN : (dir.z / dir.x); }

// C_InfiniteLineIntersection (collision.c:820) incl. the preserved "l2 vertical" y bug (:839-840),
// with the two slopes supplied by the caller
__device__ __forceinline__ bool line_isect_s(const v2 p1, const float s1, const v2 p2, const float s2, v2 &out)
{
    const bool n1 = isnan(s1), n2 = isnan(s2);
    if (n1 && n2) return false;
    if (fabsf(s1 - s2) < EPS_F) return false;
    if (n1 && !n2) {
        out.x = p1.x;
        out.z = (p1.x - p2.x) * s2 + p2.z;
    } else if (!n1 && n2) {
        out.x = p2.x;
        out.z = (p2.x - p1.x) * s1 + p2.z;
    } else {
        out.x = (s1 * p1.x - s2 * p2.x + p2.z - p1.z) / (s1 - s2);
        out.z = s2 * (out.x - p2.x) + p2.z;
    }
    return true;
}
__device__ __forceinline__ bool line_isect(const ray l1, const ray l2, v2 &out)
{
    return line_isect_s(l1.point, line_slope(l1.dir), l2.point, line_slope(l2.dir), out);
}

// (a / b < 0.0f) for |b| <= 1 without the division: the quotient of a non-zero a by such a b never
// underflows to zero, so only the signs (and NaNs, and b == +-0) matter. Bit-equivalent to the
// reference's four tests in C_RayRayIntersection2D (collision.c:861-871).
__device__ __forceinline__ bool quot_lt0(float a, float b)
{
    const bool bneg = signbit(b), bnum = (b == b);
    return bnum && ((a < 0.0f && !bneg) || (a > 0.0f && bneg));
}

// compute_vo_edges (clearpath.c:130)
__device__ __forceinline__ void vo_edges(const cp_ent ent, const cp_ent nb, v2 &right, v2 &left)
{
    v2 e2n = v2_normal(v2_sub(nb.pos, ent.pos));
    v2 r = {-e2n.z, e2n.x};
    r = v2_scale(r, nb.radius + ent.radius + 0.0f);
    const v2 rt = v2_add(nb.pos, r), lt = v2_sub(nb.pos, r);
    right = v2_normal(v2_sub(rt, ent.pos));
    left = v2_normal(v2_sub(lt, ent.pos));
}

#define VEL_WARPS_PER_CTA 4
#ifndef VEL_MIN_CTAS_SINGLE
#define VEL_MIN_CTAS_SINGLE 6
#endif
#ifndef VEL_MIN_CTAS
#define VEL_MIN_CTAS 6          // 24 warps / SM: <= 80 registers (4 B of spill in the single-pass variant), 6 x 32 KB of shared memory
#endif
#define CQ_CAP 160
struct VelSmem {
    cp_ent dyn[PFNAV_MAX_NEIGHBOURS];
    cp_ent stat[PFNAV_MAX_NEIGHBOURS];
    // rays, SoA (2 per velocity obstacle: even = left side, odd = right side)
    float rpx[4 * PFNAV_MAX_NEIGHBOURS], rpz[4 * PFNAV_MAX_NEIGHBOURS];
    float rdx[4 * PFNAV_MAX_NEIGHBOURS], rdz[4 * PFNAV_MAX_NEIGHBOURS];
    float rsl[4 * PFNAV_MAX_NEIGHBOURS];
    uint32_t near_id[128];
    float2 term[128];
    float cqx[CQ_CAP], cqz[CQ_CAP];      // candidate points awaiting the inside-PCR test
    int cqk[CQ_CAP];                      // their sequence index in the reference's push order
    // ---- one-pass emulation of the drop-furthest retry loop (clearpath_retry) ----
    uint8_t vo_src[64];                   // obstacle -> neighbour slot it was built from (dyn: 0..31, stat: 32..63)
    uint8_t nb_rank[64];                  // neighbour slot -> removal time 1.. (255: still there when the loop ends)
    uint8_t vo_rank[64];                  // obstacle -> removal time of its neighbour
    uint8_t vo_ord[64];                   // obstacles by decreasing removal time
    uint8_t cur[64];                      // working lists: dyn slots at [0, nd), stat slots at [32, 32 + ns)
    float   ndist[64];                    // neighbour slot -> distance from the entity
    uint8_t cqd[CQ_CAP];                  // death time of the queued candidates
};

// One velocity obstacle of inside_pcr (clearpath.c:252-287): is `test` strictly inside VO `i`?
// The reference normalises (test - apex) and compares a 2-D cross product with +-1/1024. Away from
// those thresholds the unnormalised cross product decides with certainty (margins of 1 % on the
// threshold versus ~1e-6 relative float error), so the sqrt and the two IEEE divisions are only paid
// in the thin band around a threshold, where the reference's exact sequence is replayed.
__device__ __forceinline__ bool vo_contains(const VelSmem &s, int i, v2 test)
{
    const float E_LO2 = (EPS_F * 0.99f) * (EPS_F * 0.99f), E_HI2 = (EPS_F * 1.01f) * (EPS_F * 1.01f);
    {   // left side: "left_of_vo" <=> det < EPS  -> not inside
        const int k = 2 * i;
        const v2 ptt = {test.x - s.rpx[k], test.z - s.rpz[k]};
        const float len2 = ptt.x * ptt.x + ptt.z * ptt.z;
        const float u = (ptt.z * s.rdx[k]) - (ptt.x * s.rdz[k]);
        const float u2 = u * u;
        bool decided = false, skip = false;
        if (len2 > E_HI2 && len2 < 1e30f) {             // |ptt| certainly >= EPS
            if (u <= 0.0f || u2 < E_LO2 * len2) { decided = true; skip = true; }        // det < EPS
            else if (u2 > E_HI2 * len2) { decided = true; skip = false; }               // det >= EPS
        }
        if (!decided) {
            const float len = sqrtf(len2);
            if (len < EPS_F) skip = true;
            else {
                const v2 n = {ptt.x / len, ptt.z / len};
                const float det = (n.z * s.rdx[k]) - (n.x * s.rdz[k]);
                skip = det < EPS_F;
            }
        }
        if (skip) return false;
    }
    {   // right side: "right_of_vo" <=> det > -EPS -> not inside
        const int k = 2 * i + 1;
        const v2 ptt = {test.x - s.rpx[k], test.z - s.rpz[k]};
        const float len2 = ptt.x * ptt.x + ptt.z * ptt.z;
        const float u = (ptt.z * s.rdx[k]) - (ptt.x * s.rdz[k]);
        const float u2 = u * u;
        bool decided = false, skip = false;
        if (len2 > E_HI2 && len2 < 1e30f) {
            if (u >= 0.0f || u2 < E_LO2 * len2) { decided = true; skip = true; }        // det > -EPS
            else if (u2 > E_HI2 * len2) { decided = true; skip = false; }               // det <= -EPS
        }
        if (!decided) {
            const float len = sqrtf(len2);
            if (len < EPS_F) skip = true;
            else {
                const v2 n = {ptt.x / len, ptt.z / len};
                const float det = (n.z * s.rdx[k]) - (n.x * s.rdz[k]);
                skip = det > -EPS_F;
            }
        }
        if (skip) return false;
    }
    return true;
}

// Drain the candidate queue: every lane keeps one candidate in flight and tests one velocity obstacle
// per iteration; a lane whose candidate is decided immediately takes the next one, so all lanes stay
// busy whatever the early-exit pattern of inside_pcr is. Keeps the per-lane first-minimum of
// compute_vnew (clearpath.c:368) on (distance, sequence index).
// The candidate a lane has in flight survives between two drains of the same solve (DrainLane): a drain that is not the
// last one of its solve returns as soon as the queue is empty and leaves the long-running candidates (the ones that pass
// obstacle after obstacle) in their lanes, where the next batch of candidates fills the idle lanes around them; only the
// final drain (`flush`) waits for the stragglers.
struct DrainLane { int busy = 0, vo = 0, myk = 0, vstart = 0; v2 myp = {0.0f, 0.0f}; };
// vstart: obstacle that swallowed this lane's previous candidate: tested first (any order is exact)

__device__ __forceinline__ void drain_candidates(const VelSmem &s, int qn, int nvo, const v2 ent_pos, const v2 des_v,
                                                 uint32_t lane, float &best, int &best_idx, v2 &best_p, int &any,
                                                 DrainLane &dl, bool flush)
{
    int next = 0;
    while (true) {
        const bool need = !dl.busy;
        const uint32_t mneed = __ballot_sync(FULL, need);
        const int avail = qn - next;
        if (need && avail > 0) {
            const int rank = __popc(mneed & ((1u << lane) - 1));
            if (rank < avail) { const int my = next + rank; dl.busy = 1; dl.vo = 0; dl.myp = {s.cqx[my], s.cqz[my]}; dl.myk = s.cqk[my]; }
        }
        next += min(__popc(mneed), avail);
        if (!flush && next >= qn) break;
        if (!__any_sync(FULL, dl.busy)) break;
        if (dl.busy) {
            bool finished = false, inside = false;
            int vidx = dl.vo + dl.vstart;
            if (vidx >= nvo) vidx -= nvo;
            if (dl.vo < nvo) inside = vo_contains(s, vidx, dl.myp);
            if (inside) { finished = true; dl.vstart = vidx; }
            else if (++dl.vo >= nvo) {
                finished = true;
                any = 1;
                const v2 curr = v2_sub(dl.myp, ent_pos);
                const float len = v2_len(v2_sub(des_v, curr));
                if (len < best || (len == best && best_idx != 0x7fffffff && dl.myk < best_idx)) { best = len; best_idx = dl.myk; best_p = curr; }
            }
            if (finished) dl.busy = 0;
        }
    }
}

// compute_all_hrvos / compute_all_vos (clearpath.c:216-247): the ray table of the velocity obstacles,
// one neighbour per lane, order-preserving compaction. Returns the number of rays (2 per obstacle).
__device__ int build_vos(VelSmem &s, const cp_ent ent, int ndyn, int nstat, uint32_t lane)
{
    // ---- compute_all_hrvos / compute_all_vos: one neighbour per lane, order-preserving compaction ----
    int n_rays = 0;
    {
        bool keep = false;
        v2 apex = {0.f, 0.f}, left = {0.f, 0.f}, right = {0.f, 0.f};
        if ((int)lane < ndyn) {
            const cp_ent nb = s.dyn[lane];
            if (!(v2_len(v2_sub(nb.pos, ent.pos)) < EPS_F)) {        // same_position (clearpath.c:123)
                keep = true;
                vo_edges(ent, nb, right, left);
                // compute_rvo / compute_hrvo (clearpath.c:161-214)
                const v2 rvo_apex = v2_add(ent.pos, v2_scale(v2_add(ent.vel, nb.vel), 0.5f));
                const v2 centerline = v2_add(left, right);
                const v2 vo_apex = v2_add(ent.pos, nb.vel);
                const float det = (centerline.x * ent.vel.z) - (centerline.z * ent.vel.x);
                apex = rvo_apex;
                if (det > EPS_F) {
                    v2 p;
                    if (line_isect(ray{rvo_apex, left}, ray{vo_apex, right}, p)) apex = p;
                } else if (det < -EPS_F) {
                    v2 p;
                    if (line_isect(ray{rvo_apex, right}, ray{vo_apex, left}, p)) apex = p;
                }
            }
        }
        const uint32_t m = __ballot_sync(FULL, keep);
        if (keep) {
            const int k = 2 * __popc(m & ((1u << lane) - 1));
            s.vo_src[k >> 1] = (uint8_t)lane;
            s.rpx[k] = apex.x; s.rpz[k] = apex.z; s.rdx[k] = left.x; s.rdz[k] = left.z; s.rsl[k] = line_slope(left);
            s.rpx[k + 1] = apex.x; s.rpz[k + 1] = apex.z; s.rdx[k + 1] = right.x; s.rdz[k + 1] = right.z; s.rsl[k + 1] = line_slope(right);
        }
        n_rays = 2 * __popc(m);
    }
    {
        bool keep = false;
        v2 apex = {0.f, 0.f}, left = {0.f, 0.f}, right = {0.f, 0.f};
        if ((int)lane < nstat) {
            const cp_ent nb = s.stat[lane];
            if (!(v2_len(v2_sub(nb.pos, ent.pos)) < EPS_F)) {
                keep = true;
                vo_edges(ent, nb, right, left);
                apex = v2_add(ent.pos, nb.vel);                       // compute_vo (clearpath.c:153)
            }
        }
        const uint32_t m = __ballot_sync(FULL, keep);
        if (keep) {
            const int k = n_rays + 2 * __popc(m & ((1u << lane) - 1));
            s.vo_src[k >> 1] = (uint8_t)(32 + lane);
            s.rpx[k] = apex.x; s.rpz[k] = apex.z; s.rdx[k] = left.x; s.rdz[k] = left.z; s.rsl[k] = line_slope(left);
            s.rpx[k + 1] = apex.x; s.rpz[k + 1] = apex.z; s.rdx[k + 1] = right.x; s.rdz[k + 1] = right.z; s.rsl[k + 1] = line_slope(right);
        }
        n_rays += 2 * __popc(m);
    }
    __syncwarp();
    return n_rays;
}

// clearpath_new_velocity (clearpath.c:552). Warp-cooperative; returns a warp-uniform status.
__device__ bool clearpath_new_velocity(VelSmem &s, const cp_ent ent, const v2 des_v, int ndyn, int nstat,
                                       uint32_t lane, v2 &out, int &n_rays_out)
{
    const int n_rays = build_vos(s, ent, ndyn, nstat, lane);
    n_rays_out = n_rays;
    const int nvo = n_rays >> 1;

    // ---- is the preferred velocity admissible? one velocity obstacle per lane ----
    const v2 des_v_ws = v2_add(ent.pos, des_v);
    bool in_any = false;
    for (int i = lane; i < nvo; i += 32) in_any |= vo_contains(s, i, des_v_ws);
    if (!__any_sync(FULL, in_any)) {
        out = des_v;
        return true;
    }

    // ---- compute_vo_xpoints + compute_vdes_proj_points + compute_vnew, fused: ray pairs are intersected
    //      32 at a time, the survivors are compacted into a small queue, and the queue is drained by
    //      drain_candidates(); the candidate list is never materialised ----
    float best = __int_as_float(0x7f800000);    // +inf ; `len < min_dist` with min_dist = INFINITY
    int best_idx = 0x7fffffff;
    v2 best_p = {0.0f, 0.0f};
    int any = 0, qn = 0;
    DrainLane dl;
    const int npairs = n_rays * n_rays;
    // Exact pruning: compute_vnew (clearpath.c:368) keeps the admissible candidate nearest to des_v (first
    // one on ties), so a candidate that is certainly farther than an admissible candidate already found
    // can neither win nor change the tie-break and need not be tested against the obstacles at all. The
    // projection points of compute_vdes_proj_points (sequence indices npairs..npairs+n_rays-1) are the
    // nearest points of each ray to des_v: they are drained FIRST to get a tight bound, then the ray-pair
    // intersections are filtered against the warp-wide bound before they enter the queue.
    for (int base = 0; base < n_rays; base += 32) {
        const int r = base + (int)lane;
        const bool ok = r < n_rays;
        if (ok) {
            const v2 d = {s.rdx[r], s.rdz[r]};
            const float len = v2_dot(d, des_v);
            const v2 p = v2_add(v2{s.rpx[r], s.rpz[r]}, v2_scale(d, len));
            s.cqx[qn + (int)lane] = p.x; s.cqz[qn + (int)lane] = p.z; s.cqk[qn + (int)lane] = npairs + r;
        }
        qn += min(32, n_rays - base);
        if (qn > CQ_CAP - 32 || base + 32 >= n_rays) {
            __syncwarp();
            drain_candidates(s, qn, nvo, ent.pos, des_v, lane, best, best_idx, best_p, any, dl, true);     // the bound below needs them all
            __syncwarp();
            qn = 0;
        }
    }
    float bound2 = __int_as_float(0x7f800000);
    auto refresh_bound = [&]() {
        float b = best;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) b = fminf(b, __shfl_xor_sync(FULL, b, off));
        bound2 = b * b * 1.00001f;              // margin >> float rounding of the two squared lengths
    };
    refresh_bound();
    int i = 0, j = (int)lane;                   // pair index of this lane: k = i * n_rays + j
    while (j >= n_rays) { j -= n_rays; i++; }
    for (int base = 0; base < npairs; base += 32) {
        const int k = base + (int)lane;
        v2 p = {0.f, 0.f};
        bool ok = false;
        if (k < npairs) {
            if (i != j) {
                // C_RayRayIntersection2D (collision.c:854)
                const v2 p1 = {s.rpx[i], s.rpz[i]}, p2 = {s.rpx[j], s.rpz[j]};
                if (line_isect_s(p1, s.rsl[i], p2, s.rsl[j], p)) {
                    const float ddx = des_v.x - (p.x - ent.pos.x), ddz = des_v.z - (p.z - ent.pos.z);
                    ok = !(ddx * ddx + ddz * ddz > bound2) &&
                         !(quot_lt0(p.x - p1.x, s.rdx[i]) || quot_lt0(p.z - p1.z, s.rdz[i]) ||
                           quot_lt0(p.x - p2.x, s.rdx[j]) || quot_lt0(p.z - p2.z, s.rdz[j]));
                }
            }
        }
        const uint32_t m = __ballot_sync(FULL, ok);
        if (ok) {
            const int q = qn + __popc(m & ((1u << lane) - 1));
            s.cqx[q] = p.x; s.cqz[q] = p.z; s.cqk[q] = k;
        }
        qn += __popc(m);
        j += 32;
        while (j >= n_rays) { j -= n_rays; i++; }
        if (qn > CQ_CAP - 32 || base + 32 >= npairs) {
            __syncwarp();
            drain_candidates(s, qn, nvo, ent.pos, des_v, lane, best, best_idx, best_p, any, dl, base + 32 >= npairs);
            __syncwarp();
            qn = 0;
            refresh_bound();
        }
    }
    any = __any_sync(FULL, any);
    if (!any) return false;
    // warp arg-min on (len, idx); lanes that never improved hold (+inf, INT_MAX)
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const float ob = __shfl_xor_sync(FULL, best, off);
        const int oi = __shfl_xor_sync(FULL, best_idx, off);
        const float ox = __shfl_xor_sync(FULL, best_p.x, off), oz = __shfl_xor_sync(FULL, best_p.z, off);
        if (ob < best || (ob == best && oi < best_idx)) { best = ob; best_idx = oi; best_p = {ox, oz}; }
    }
    out = (best_idx == 0x7fffffff) ? v2{0.0f, 0.0f} : best_p;
    return true;
}

// ------------------------------------------------------------------------------------------
// G_ClearPath_NewVelocity's retry loop (clearpath.c:702-713) in ONE pass. After a solve that finds no admissible
// velocity the reference drops the furthest neighbour (remove_furthest :390) and solves again from scratch, while both
// neighbour lists are non-empty. Which neighbour goes at which step depends on the distances and the list order only,
// never on the velocities, so the whole removal schedule is known up front: neighbour n leaves at time rank(n). The
// obstacle of a neighbour and the candidate points generated by its two rays do not depend on the other neighbours,
// hence for every candidate point c of the FIRST solve
//     adm(c)   = max rank of the obstacles that contain c   (it becomes admissible once they are all gone)
//     death(c) = min rank of the obstacles whose rays generate c
// and the solve after t removals succeeds iff the preferred velocity (adm <= t) or some candidate (adm <= t < death) is
// admissible. The first such t is T = min(adm); by minimality every candidate usable at T has adm == T exactly, and
// compute_vnew (:368) picks the one nearest to the preferred velocity, first in ITS list order on ties -- the order of
// the ray table after T swap-with-last deletions; a tie between two different points is therefore flagged and that (rare)
// entity replays the loop literally. Obstacles are tested in decreasing rank, so
// the first hit is the maximum and most points end at the nearest (widest) obstacles after one or two tests.
// Cost: one solve instead of up to 63.
// ------------------------------------------------------------------------------------------
// obstacle ranks from the neighbour ranks, and the obstacles ordered by decreasing rank (index ascending among equals)
__device__ __forceinline__ void retry_rank_vos(VelSmem &s, int nvo, uint32_t lane)
{
    for (int v = lane; v < nvo; v += 32) s.vo_rank[v] = s.nb_rank[s.vo_src[v]];
    __syncwarp();
    for (int v = lane; v < nvo; v += 32) {
        const int r = s.vo_rank[v];
        int p = 0;
        for (int u = 0; u < nvo; u++) { const int ru = s.vo_rank[u]; p += (ru > r) || (ru == r && u < v); }
        s.vo_ord[p] = (uint8_t)v;
    }
    __syncwarp();
}

// removal schedule; returns the number of removals after which the loop ends (no solve happens at that time)
__device__ int retry_schedule(VelSmem &s, const v2 pos, int ndyn, int nstat, int nvo, uint32_t lane)
{
    for (int k = lane; k < 64; k += 32) {
        const bool valid = k < 32 ? k < ndyn : (k - 32) < nstat;
        float d = -__int_as_float(0x7f800000);
        if (valid) { const cp_ent e = k < 32 ? s.dyn[k] : s.stat[k - 32]; d = v2_len(v2_sub(pos, e.pos)); }
        s.ndist[k] = d; s.nb_rank[k] = 255; s.cur[k] = (uint8_t)k;
    }
    __syncwarp();
    int nd = ndyn, ns = nstat, t = 0, t_end = 0;
    while (true) {
        float bd = -__int_as_float(0x7f800000);
        int bp = 0x7fffffff;
        for (int q = lane; q < nd + ns; q += 32) {
            const int slot = q < nd ? s.cur[q] : s.cur[32 + q - nd];
            const float d = s.ndist[slot];
            if (d > bd) { bd = d; bp = q; }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const float od = __shfl_xor_sync(FULL, bd, off);
            const int op = __shfl_xor_sync(FULL, bp, off);
            if (od > bd || (od == bd && op < bp)) { bd = od; bp = op; }
        }
        if (bp == 0x7fffffff) { t_end = t + 1; break; }      // nothing left to remove: the loop condition fails next
        t++;
        if (lane == 0) {
            int slot;
            if (bp < nd) { slot = s.cur[bp]; s.cur[bp] = s.cur[nd - 1]; }
            else { const int j = bp - nd; slot = s.cur[32 + j]; s.cur[32 + j] = s.cur[32 + ns - 1]; }
            s.nb_rank[slot] = (uint8_t)t;
        }
        if (bp < nd) nd--; else ns--;
        __syncwarp();
        if (!(nd > 0 && ns > 0)) { t_end = t; break; }
    }
    retry_rank_vos(s, nvo, lane);
    return t_end;
}

// The same schedule without the serial loop, for the common case that no two neighbours are exactly equally far away:
// the loop then removes them in descending distance whatever the list order, neighbour n leaves at time
// 1 + #{m : d(m) > d(n)}, and the loop ends when the last member of either list has left. Equal distances (lattice
// crowds) fall back to the literal schedule above, whose tie-break follows the swap-with-last list order.
__device__ int retry_schedule_fast(VelSmem &s, const v2 pos, int ndyn, int nstat, int nvo, uint32_t lane)
{
    for (int k = lane; k < 64; k += 32) {
        const bool valid = k < 32 ? k < ndyn : (k - 32) < nstat;
        float d = -__int_as_float(0x7f800000);
        if (valid) { const cp_ent e = k < 32 ? s.dyn[k] : s.stat[k - 32]; d = v2_len(v2_sub(pos, e.pos)); }
        s.ndist[k] = d;
    }
    __syncwarp();
    bool tie = false;
    int rk[2] = {0, 0};
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int k = (int)lane + 32 * h;
        const bool valid = h == 0 ? (int)lane < ndyn : (int)lane < nstat;
        if (valid) {
            const float d = s.ndist[k];
            int r = 1;
            for (int j = 0; j < ndyn; j++) { const float dj = s.ndist[j]; r += dj > d; tie |= (dj == d && j != k); }
            for (int j = 0; j < nstat; j++) { const float dj = s.ndist[32 + j]; r += dj > d; tie |= (dj == d && 32 + j != k); }
            rk[h] = r;
        }
    }
    if (__any_sync(FULL, tie)) return retry_schedule(s, pos, ndyn, nstat, nvo, lane);
    int md = rk[0], ms = rk[1];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) { md = max(md, __shfl_xor_sync(FULL, md, off)); ms = max(ms, __shfl_xor_sync(FULL, ms, off)); }
    const int t_end = min(md, ms);
    s.nb_rank[lane] = (uint8_t)(((int)lane < ndyn && rk[0] <= t_end) ? rk[0] : 255);
    s.nb_rank[32 + lane] = (uint8_t)(((int)lane < nstat && rk[1] <= t_end) ? rk[1] : 255);
    __syncwarp();
    retry_rank_vos(s, nvo, lane);
    return t_end;
}

struct retry_best { int T; float dist; int id; v2 p; int tie; };

// (time, distance) order. Two DIFFERENT points at exactly the same time and distance would be ordered by their position in
// the candidate list of the solve after T removals (compute_vnew keeps the first, clearpath.c:368) -- that list order
// depends on the swap-with-last deletions done so far; instead of reproducing it the tie is flagged and the caller
// replays the reference's loop literally (it practically never happens: identical points, e.g. xpoint(i, j) and
// xpoint(j, i), are no tie -- either one gives the same velocity).
__device__ __forceinline__ bool retry_better(int T, float dist, const v2 p, retry_best &b)
{
    if (T != b.T) return T < b.T;
    if (dist != b.dist) return dist < b.dist;
    if (b.id >= 0 && (p.x != b.p.x || p.z != b.p.z)) b.tie = 1;
    return false;
}

// the outcome of the retry loop from the per-lane bests: smallest time, then distance, then list position at that time
__device__ __forceinline__ v2 retry_finish(retry_best &best, int adm_des, int t_end, const v2 des_v, uint32_t lane, bool &exact)
{
    int T = best.T;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) T = min(T, __shfl_xor_sync(FULL, T, off));
    if (best.T != T) { best.dist = __int_as_float(0x7f800000); best.id = -1; best.tie = 0; }
    if (adm_des <= T && adm_des <= t_end - 1) return des_v;     // inside_pcr(des_v) is tested before any candidate (:602)
    float d = best.dist;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) d = fminf(d, __shfl_xor_sync(FULL, d, off));
    const bool mine = best.id >= 0 && best.dist == d;
    const uint32_t mm = __ballot_sync(FULL, mine);
    if (!mm) return v2{0.0f, 0.0f};                              // no solve before the loop ends finds a point
    const int src = __ffs(mm) - 1;
    v2 out;
    out.x = __shfl_sync(FULL, best.p.x, src); out.z = __shfl_sync(FULL, best.p.z, src);
    // ties between different points, inside a lane or across lanes: list order would decide
    const bool differs = mine && (best.tie || best.p.x != out.x || best.p.z != out.z);
    if (__any_sync(FULL, differs)) exact = false;
    return out;
}

// like drain_candidates, for the retry emulation: cqk holds the candidate id, cqd its death time
__device__ __forceinline__ void drain_ranked(const VelSmem &s, int qn, int nvo, const v2 ent_pos, const v2 des_v,
                                             uint32_t lane, retry_best &best)
{
    int next = 0, my = -1, k = 0, myid = 0, mydeath = 0;
    v2 myp = {0.0f, 0.0f};
    while (true) {
        const bool need = my < 0;
        const uint32_t mneed = __ballot_sync(FULL, need);
        const int avail = qn - next;
        if (need) {
            const int rank = __popc(mneed & ((1u << lane) - 1));
            if (rank < avail) { my = next + rank; k = 0; myp = {s.cqx[my], s.cqz[my]}; myid = s.cqk[my]; mydeath = s.cqd[my]; }
        }
        next += min(__popc(mneed), avail);
        if (!__any_sync(FULL, my >= 0)) break;
        if (my >= 0) {
            bool finished = false;
            int adm = -1;
            if (k < nvo) {
                const int v = s.vo_ord[k];
                const int r = s.vo_rank[v];
                // obstacles come in decreasing rank, so the first hit is the maximum rank among the containing ones
                if (vo_contains(s, v, myp)) { adm = r; finished = true; }
                else if (++k >= nvo) { adm = 0; finished = true; }
            } else { adm = 0; finished = true; }
            if (finished) {
                if (adm < mydeath && adm <= best.T) {
                    const v2 curr = v2_sub(myp, ent_pos);
                    const float len = v2_len(v2_sub(des_v, curr));
                    if (retry_better(adm, len, curr, best)) { best.T = adm; best.dist = len; best.id = myid; best.p = curr; best.tie = 0; }
                }
                my = -1;
            }
        }
    }
}

// everything G_ClearPath_NewVelocity does after its first solve found nothing. The ray table of that solve is still in
// shared memory (build_vos). Returns the velocity of the first successful later solve, or zero when the loop ends first;
// `exact` is cleared when an order-dependent tie was met (the caller then replays the loop literally).
__device__ v2 clearpath_retry(VelSmem &s, const cp_ent ent, const v2 des_v, int ndyn, int nstat, int n_rays, uint32_t lane, bool &exact)
{
    exact = true;
    const int nvo = n_rays >> 1;
    // with one of the two lists empty the loop condition fails right after the first removal (the common case in a crowd
    // where everybody moves): no second solve, the answer is zero
    if (ndyn == 0 || nstat == 0) return v2{0.0f, 0.0f};
    const int t_end = retry_schedule_fast(s, ent.pos, ndyn, nstat, nvo, lane);
    if (t_end <= 1) return v2{0.0f, 0.0f};                   // the loop ends right after the first removal
    // the preferred velocity: admissible once every obstacle that contains it is gone
    const v2 des_v_ws = v2_add(ent.pos, des_v);
    int adm_des = 0;
    for (int v = lane; v < nvo; v += 32)
        if (vo_contains(s, v, des_v_ws)) adm_des = max(adm_des, (int)s.vo_rank[v]);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) adm_des = max(adm_des, __shfl_xor_sync(FULL, adm_des, off));
    retry_best best; best.T = min(adm_des, t_end - 1); best.dist = __int_as_float(0x7f800000); best.id = -1; best.p = {0.0f, 0.0f}; best.tie = 0;
    const int npairs = n_rays * n_rays;
    int qn = 0;
    auto flush = [&](bool last) {
        if (qn > CQ_CAP - 32 || (last && qn > 0)) {
            __syncwarp();
            drain_ranked(s, qn, nvo, ent.pos, des_v, lane, best);
            __syncwarp();
            qn = 0;
            int t = best.T;                                   // the warp-wide best time bounds what is still worth testing
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) t = min(t, __shfl_xor_sync(FULL, t, off));
            if (t < best.T) { best.T = t; best.dist = __int_as_float(0x7f800000); best.id = -1; best.tie = 0; }
        }
    };
    for (int base = 0; base < n_rays; base += 32) {            // projection points (compute_vdes_proj_points, :344)
        const int r = base + (int)lane;
        if (r < n_rays) {
            const v2 d = {s.rdx[r], s.rdz[r]};
            const float len = v2_dot(d, des_v);
            const v2 p = v2_add(v2{s.rpx[r], s.rpz[r]}, v2_scale(d, len));
            s.cqx[qn + (int)lane] = p.x; s.cqz[qn + (int)lane] = p.z; s.cqk[qn + (int)lane] = 0x8000 | r;
            s.cqd[qn + (int)lane] = s.vo_rank[r >> 1];
        }
        qn += min(32, n_rays - base);
        flush(base + 32 >= n_rays);
    }
    int i = 0, j = (int)lane;
    while (j >= n_rays && n_rays > 0) { j -= n_rays; i++; }
    for (int base = 0; base < npairs; base += 32) {            // ray-pair intersections (compute_vo_xpoints, :321)
        const int k = base + (int)lane;
        v2 p = {0.f, 0.f};
        bool ok = false;
        int death = 0;
        if (k < npairs && i != j) {
            death = min((int)s.vo_rank[i >> 1], (int)s.vo_rank[j >> 1]);
            if (death > 0) {           // a pair that dies at time <= 0 cannot exist (ranks start at 1); kept for symmetry
                const v2 p1 = {s.rpx[i], s.rpz[i]}, p2 = {s.rpx[j], s.rpz[j]};
                if (line_isect_s(p1, s.rsl[i], p2, s.rsl[j], p))
                    ok = !(quot_lt0(p.x - p1.x, s.rdx[i]) || quot_lt0(p.z - p1.z, s.rdz[i]) ||
                           quot_lt0(p.x - p2.x, s.rdx[j]) || quot_lt0(p.z - p2.z, s.rdz[j]));
            }
        }
        const uint32_t m = __ballot_sync(FULL, ok);
        if (ok) {
            const int q = qn + __popc(m & ((1u << lane) - 1));
            s.cqx[q] = p.x; s.cqz[q] = p.z; s.cqk[q] = (i << 7) | j; s.cqd[q] = (uint8_t)death;
        }
        qn += __popc(m);
        j += 32;
        while (j >= n_rays) { j -= n_rays; i++; }
        flush(base + 32 >= npairs);
    }
    return retry_finish(best, adm_des, t_end, des_v, lane, exact);
}

// remove_furthest (clearpath.c:390): first strict maximum over dyn then stat; swap-with-last delete
__device__ void remove_furthest(VelSmem &s, const v2 pos, int &ndyn, int &nstat, uint32_t lane)
{
    float d = -__int_as_float(0x7f800000);
    int idx = 0x7fffffff;
    const int n = ndyn + nstat;
    // n <= 64: two candidates per lane, sequence index = position in (dyn ++ stat)
    for (int k = lane; k < n; k += 32) {
        const cp_ent e = k < ndyn ? s.dyn[k] : s.stat[k - ndyn];
        const float len = v2_len(v2_sub(pos, e.pos));
        if (len > d) { d = len; idx = k; }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const float od = __shfl_xor_sync(FULL, d, off);
        const int oi = __shfl_xor_sync(FULL, idx, off);
        if (od > d || (od == d && oi < idx)) { d = od; idx = oi; }
    }
    if (idx != 0x7fffffff) {
        if (lane == 0) {
            if (idx < ndyn) s.dyn[idx] = s.dyn[ndyn - 1];
            else s.stat[idx - ndyn] = s.stat[nstat - 1];
        }
        if (idx < ndyn) ndyn--; else nstat--;
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------
// Two-phase ClearPath. Everything G_ClearPath_NewVelocity does except the final choice is independent of
// the preferred velocity: the velocity obstacles depend on positions / velocities only, and so do the
// pairwise ray intersections (compute_vo_xpoints, clearpath.c:321) and their inside-PCR tests. Phase A
// (k_agent_velocity<1>, runs while the LOS chains are still in flight) therefore collects the ADMISSIBLE
// intersection points of every agent; phase B (k_agent_velocity<2>, after the fields are joined) computes
// the preferred velocity, tests it and its projections (compute_vdes_proj_points, :344) and picks the
// nearest admissible point by (distance, sequence index) = compute_vnew's first minimum (:368).
// ------------------------------------------------------------------------------------------
struct pf_xpoint { float x, z; int k; };
#define PF_PREP_XP_CAP 64
#define PF_PREP_OVERFLOW 0xFFFFFFFFu
struct pf_prep {
    uint32_t ndyn, nstat, nx, _pad;
    float sepx, sepz;
    uint32_t dyn_id[PFNAV_MAX_NEIGHBOURS], stat_id[PFNAV_MAX_NEIGHBOURS];
    pf_xpoint xp[PF_PREP_XP_CAP];
};

// like drain_candidates, but records every candidate that lies inside no velocity obstacle
__device__ __forceinline__ void drain_collect(const VelSmem &s, int qn, int nvo, uint32_t lane, pf_xpoint *out, int &cnt,
                                              DrainLane &dl, bool flush)
{
    int next = 0;
    while (true) {
        const bool need = !dl.busy;
        const uint32_t mneed = __ballot_sync(FULL, need);
        const int avail = qn - next;
        if (need && avail > 0) {
            const int rank = __popc(mneed & ((1u << lane) - 1));
            if (rank < avail) { const int my = next + rank; dl.busy = 1; dl.vo = 0; dl.myp = {s.cqx[my], s.cqz[my]}; dl.myk = s.cqk[my]; }
        }
        next += min(__popc(mneed), avail);
        if (!flush && next >= qn) break;
        if (!__any_sync(FULL, dl.busy)) break;
        bool admissible = false;
        if (dl.busy) {
            int vidx = dl.vo + dl.vstart;
            if (vidx >= nvo) vidx -= nvo;
            bool inside = false;
            if (dl.vo < nvo) inside = vo_contains(s, vidx, dl.myp);
            if (inside) { dl.busy = 0; dl.vstart = vidx; }
            else if (++dl.vo >= nvo) { admissible = true; dl.busy = 0; }
        }
        const uint32_t ma = __ballot_sync(FULL, admissible);
        if (admissible) {
            const int q = cnt + __popc(ma & ((1u << lane) - 1));
            if (q < PF_PREP_XP_CAP) { out[q].x = dl.myp.x; out[q].z = dl.myp.z; out[q].k = dl.myk; }
        }
        cnt += __popc(ma);
    }
}

// phase A: admissible ray-pair intersections of one agent -> out[0..cnt) (cnt > cap = overflow)
__device__ int clearpath_collect(VelSmem &s, const cp_ent ent, int ndyn, int nstat, uint32_t lane, pf_xpoint *out)
{
    const int n_rays = build_vos(s, ent, ndyn, nstat, lane);
    const int nvo = n_rays >> 1, npairs = n_rays * n_rays;
    int cnt = 0, qn = 0;
    DrainLane dl;
    int i = 0, j = (int)lane;
    while (j >= n_rays && n_rays > 0) { j -= n_rays; i++; }
    for (int base = 0; base < npairs; base += 32) {
        const int k = base + (int)lane;
        v2 p = {0.f, 0.f};
        bool ok = false;
        if (k < npairs && i != j) {
            const v2 p1 = {s.rpx[i], s.rpz[i]}, p2 = {s.rpx[j], s.rpz[j]};
            if (line_isect_s(p1, s.rsl[i], p2, s.rsl[j], p)) {
                ok = !(quot_lt0(p.x - p1.x, s.rdx[i]) || quot_lt0(p.z - p1.z, s.rdz[i]) ||
                       quot_lt0(p.x - p2.x, s.rdx[j]) || quot_lt0(p.z - p2.z, s.rdz[j]));
            }
        }
        const uint32_t m = __ballot_sync(FULL, ok);
        if (ok) {
            const int q = qn + __popc(m & ((1u << lane) - 1));
            s.cqx[q] = p.x; s.cqz[q] = p.z; s.cqk[q] = k;
        }
        qn += __popc(m);
        j += 32;
        while (j >= n_rays) { j -= n_rays; i++; }
        if (qn > CQ_CAP - 32 || base + 32 >= npairs) {
            __syncwarp();
            drain_collect(s, qn, nvo, lane, out, cnt, dl, base + 32 >= npairs);
            __syncwarp();
            qn = 0;
        }
    }
    return cnt;
}

// phase B: the choice among {des_v, its projections on the rays, the stored admissible intersections}
__device__ bool clearpath_finish(VelSmem &s, const cp_ent ent, const v2 des_v, int ndyn, int nstat, uint32_t lane,
                                 const pf_xpoint *xp, int nx, v2 &out, int &n_rays_out)
{
    const int n_rays = build_vos(s, ent, ndyn, nstat, lane);
    n_rays_out = n_rays;
    const int nvo = n_rays >> 1, npairs = n_rays * n_rays;
    const v2 des_v_ws = v2_add(ent.pos, des_v);
    bool in_any = false;
    for (int i = lane; i < nvo; i += 32) in_any |= vo_contains(s, i, des_v_ws);
    if (!__any_sync(FULL, in_any)) { out = des_v; return true; }
    float best = __int_as_float(0x7f800000);
    int best_idx = 0x7fffffff;
    v2 best_p = {0.0f, 0.0f};
    int any = 0, qn = 0;
    for (int base = 0; base < n_rays; base += 32) {
        const int r = base + (int)lane;
        if (r < n_rays) {
            const v2 d = {s.rdx[r], s.rdz[r]};
            const float len = v2_dot(d, des_v);
            const v2 p = v2_add(v2{s.rpx[r], s.rpz[r]}, v2_scale(d, len));
            s.cqx[qn + (int)lane] = p.x; s.cqz[qn + (int)lane] = p.z; s.cqk[qn + (int)lane] = npairs + r;
        }
        qn += min(32, n_rays - base);
        if (qn > CQ_CAP - 32 || base + 32 >= n_rays) {
            __syncwarp();
            DrainLane dl;
            drain_candidates(s, qn, nvo, ent.pos, des_v, lane, best, best_idx, best_p, any, dl, true);
            __syncwarp();
            qn = 0;
        }
    }
    for (int q = lane; q < nx; q += 32) {       // the intersections phase A found admissible
        const pf_xpoint c = xp[q];
        const v2 curr = v2_sub(v2{c.x, c.z}, ent.pos);
        const float len = v2_len(v2_sub(des_v, curr));
        any = 1;
        if (len < best || (len == best && best_idx != 0x7fffffff && c.k < best_idx)) { best = len; best_idx = c.k; best_p = curr; }
    }
    any = __any_sync(FULL, any);
    if (!any) return false;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const float ob = __shfl_xor_sync(FULL, best, off);
        const int oi = __shfl_xor_sync(FULL, best_idx, off);
        const float ox = __shfl_xor_sync(FULL, best_p.x, off), oz = __shfl_xor_sync(FULL, best_p.z, off);
        if (ob < best || (ob == best && oi < best_idx)) { best = ob; best_idx = oi; best_p = {ox, oz}; }
    }
    out = (best_idx == 0x7fffffff) ? v2{0.0f, 0.0f} : best_p;
    return true;
}

struct TickParams {
    int hz;
    float scaled_max_force;       // (float)SCALED_MAX_FORCE, as passed to vec2_truncate
    double scaled_max_force_d;    // SCALED_MAX_FORCE as the double it is in comparisons (movement.c:1895)
    unsigned long long *stats;    // ClearPath event counters (pfnav_agents_clearpath_stats), never read by the kernels
    unsigned int *queue;          // next work item of the velocity kernels, one counter per MODE; zeroed every tick
};

// ------------------------------------------------------------------------------------------
// K6c: one warp per agent
// ------------------------------------------------------------------------------------------
// MODE 0: the whole update in one pass. MODE 1 / 2: phase A / phase B of the two-phase scheme above.
template <int MODE>
__global__ void __launch_bounds__(VEL_WARPS_PER_CTA * 32, MODE == 0 ? VEL_MIN_CTAS_SINGLE : VEL_MIN_CTAS)
k_agent_velocity(MapView m, GridView g, TickParams tp, const pfnav_agent *__restrict__ agents,
                 const pf_record *__restrict__ rec, const pfnav_flock *__restrict__ flocks,
                 const uint32_t *__restrict__ work, int nwork, const float2 *__restrict__ vdes_in,
                 const uint8_t *__restrict__ los_in, const float2 *__restrict__ cohesion_in,
                 float2 *__restrict__ vel_out, float2 *__restrict__ vpref_out, int filter_garr,
                 uint32_t *__restrict__ nb_scratch, pf_prep *__restrict__ prep,
                 const pfnav_formation_in *__restrict__ form)
{
    extern __shared__ __align__(16) uint8_t vel_smem_raw[];      // VEL_WARPS_PER_CTA x VelSmem (> 48 KB: opt-in, pfnav_agents_init)
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    VelSmem &s = reinterpret_cast<VelSmem *>(vel_smem_raw)[warp];
    // work items are taken from a queue, not strided: an entity whose ClearPath solve runs long (no admissible velocity, a
    // sixth of a moving crowd) would otherwise hold its CTA's three other warps idle at the end of their own strides
    for (;;) {
        int w = 0;
        if (lane == 0) w = (int)atomicAdd(tp.queue + MODE, 1u);
        w = __shfl_sync(FULL, w, 0);
        if (w >= nwork) break;
        const uint32_t uid = work[w];
        const pfnav_agent a = agents[uid];
        const uint32_t ent_flags = a.flags & 0xFFFFFFu;
        // COMBAT_HELD (movement.c:3405)
        if (ent_flags & PFNAV_FLAG_COMBAT_HELD) {
            if (MODE != 1 && lane == 0) { vel_out[w] = make_float2(0.f, 0.f); vpref_out[w] = make_float2(0.f, 0.f); }
            continue;
        }
        const v2 pos = {a.pos[0], a.pos[1]};
        const v2 velocity = {a.velocity[0], a.velocity[1]};
        const float2 vd = MODE == 1 ? make_float2(0.f, 0.f) : vdes_in[w];      // phase A runs before the fields exist
        const v2 vdes = {vd.x, vd.y};
        const bool has_los = MODE == 1 ? false : los_in[w] != 0;
        const float hzf = (float)tp.hz;

        v2 vpref = {0.0f, 0.0f};
        if (a.state == PFNAV_STATE_TURNING) {
            vpref = {0.0f, 0.0f};
        } else {
            // ================= point_seek_vpref (movement.c:1870) =================
            // ---- separation_force (movement.c:1690): 30-wu query, first 128 hits in index order ----
            v2 separation = {0.0f, 0.0f};
            if (MODE == 2) {
                separation = {prep[w].sepx, prep[w].sepz};
            } else {
            int num_near = 0;
            grid_query(g, pos.x, pos.z, 30.0f, lane, [&](bool hit, uint32_t id) -> bool {
                const uint32_t mk = __ballot_sync(FULL, hit);
                if (!mk) return false;
                const int rank = __popc(mk & ((1u << lane) - 1));
                if (hit && num_near + rank < 128) s.near_id[num_near + rank] = id;
                num_near += __popc(mk);
                return num_near >= 128;
            });
            num_near = min(num_near, 128);
            __syncwarp();
            if (filter_garr) num_near = filter_garrisoned(s.near_id, num_near, rec, lane);
            for (int k = lane; k < num_near; k += 32) {
                const uint32_t cu = s.near_id[k];
                float2 term = make_float2(0.f, 0.f);
                if (cu != uid) {
                    const pf_record r = rec[cu];
                    const uint32_t fl = r.state_flags & 0xFFFFFFu;
                    if ((fl & PFNAV_FLAG_MOVABLE) && ((ent_flags & PFNAV_FLAG_AIR) == (fl & PFNAV_FLAG_AIR))) {
                        v2 diff = {r.px - pos.x, r.pz - pos.z};
                        const float radius = a.radius + r.radius + 0.0f;
                        const float len = v2_len(diff);
                        if (!(len < EPS_F)) {
                            const float t = (len - radius * 0.85f) / len;
                            const float scale = (float)exp((double)fminf(-20.0f * t, 40.0f));
                            diff = v2_scale(diff, scale);
                            term = make_float2(diff.x, diff.z);
                        }
                    }
                }
                s.term[k] = term;
            }
            __syncwarp();
            separation = {0.0f, 0.0f};
            for (int k = 0; k < num_near; k++) {         // the reference's summation order
                const float2 t = s.term[k];
                separation.x += t.x;
                separation.z += t.y;
            }
            if (num_near == 0) separation = {0.0f, 0.0f};
            else {
                separation = v2_scale(separation, -1.0f);
                separation = v2_truncate(separation, tp.scaled_max_force);
            }
                if (MODE == 1 && lane == 0) { prep[w].sepx = separation.x; prep[w].sepz = separation.z; }
            }
            if (MODE == 1) { /* the rest of point_seek_vpref needs the desired velocity: phase B */ } else {
            // the state picks the steering variant (move_velocity_work, movement.c:3414-3448)
            const uint32_t st_ = a.state;
            pfnav_formation_in fin;
            const bool has_form = form != nullptr && (st_ == PFNAV_STATE_ARRIVING_TO_CELL || st_ == PFNAV_STATE_MOVING_IN_FORMATION);
            if (has_form) fin = form[uid];
            const bool cell_mode = st_ == PFNAV_STATE_ARRIVING_TO_CELL, form_mode = st_ == PFNAV_STATE_MOVING_IN_FORMATION;
            const bool enemy_mode = st_ == PFNAV_STATE_SEEK_ENEMIES;
            const bool ready = has_form && (fin.flags & PFNAV_FORM_ASSIGNMENT_READY);
            v2 arrive;
            if (cell_mode) {
                // ---- arrive_force_cell (movement.c:1571): NOT relative to the current velocity, not truncated ----
                const v2 cell = has_form ? v2{fin.cell_pos[0], fin.cell_pos[1]} : pos;
                v2 desired = v2_sub(cell, pos);
                const float distance = v2_len(desired);
                if (distance < 10.0f) desired = v2_scale(desired, distance / 10.0f);
                else desired = v2_scale(vdes, a.max_speed / hzf);
                arrive = desired;
            } else if (enemy_mode) {
                // ---- arrive_force_enemies (movement.c:1593) ----
                const v2 desired = v2_scale(vdes, a.max_speed / hzf);
                arrive = v2_truncate(v2_sub(desired, velocity), tp.scaled_max_force);
            } else {
            // ---- arrive_force_point (movement.c:1546) ----
            const v2 target = a.flock >= 0 ? v2{flocks[a.flock].target[0], flocks[a.flock].target[1]} : pos;
            v2 desired;
            if (has_los) {
                desired = v2_sub(target, pos);
                const float distance = v2_len(desired);
                desired = v2_normal(desired);
                desired = v2_scale(desired, a.max_speed / hzf);
                if (distance < 10.0f) desired = v2_scale(desired, distance / 10.0f);
            } else {
                desired = v2_scale(vdes, a.max_speed / hzf);
            }
            arrive = v2_truncate(v2_sub(desired, velocity), tp.scaled_max_force);
            }
            // ---- cohesion: the flock-wide pre-pass, or the formation's own forces (fstate, movement.c:215-225) ----
            const float2 ch = cohesion_in[uid];
            const v2 cohesion = (cell_mode || form_mode) ? (has_form ? v2{fin.cohesion[0], fin.cohesion[1]} : v2{0.0f, 0.0f}) : v2{ch.x, ch.y};
            const v2 alignment = has_form ? v2{fin.align[0], fin.align[1]} : v2{0.0f, 0.0f};

            const int layer = nav_layer_for(ent_flags, a.radius);
            bool on_blocked, dummy, lp, lb, rp, rb, tpth, tb, bp, bb;
            probe_tile(m, layer, pos.x, pos.z, dummy, on_blocked);
            probe_tile(m, layer, pos.x + 4.0f, pos.z, lp, lb);      // left  = x + nt_dims.x
            probe_tile(m, layer, pos.x - 4.0f, pos.z, rp, rb);      // right
            probe_tile(m, layer, pos.x, pos.z + 4.0f, tpth, tb);    // top   = z + nt_dims.z
            probe_tile(m, layer, pos.x, pos.z - 4.0f, bp, bb);      // bot

            v2 steer = {0.0f, 0.0f};
            for (int prio = 0; prio < (enemy_mode ? 1 : 3); prio++) {
                if (prio == 0) {
                    // point_seek_total_force / formation_point_seek_total_force / cell_seek_total_force /
                    // enemy_seek_total_force (movement.c:1745, 1960, 1769, 1801)
                    const v2 A = v2_scale(arrive, 0.5f), C = v2_scale(cohesion, 0.15f), S = v2_scale(separation, 0.6f);
                    const v2 AL = v2_scale(alignment, 0.15f);
                    v2 ret = {0.0f, 0.0f};
                    ret = v2_add(ret, A);
                    ret = v2_add(ret, S);
                    if (cell_mode) {
                        const v2 cell = has_form ? v2{fin.cell_pos[0], fin.cell_pos[1]} : pos;
                        if (v2_len(v2_sub(cell, pos)) > 30.0f) { ret = v2_add(ret, C); ret = v2_add(ret, AL); }     // CELL_ARRIVAL_RADIUS
                    } else if (!enemy_mode) ret = v2_add(ret, C);
                    steer = v2_truncate(ret, tp.scaled_max_force);
                } else if (prio == 1) steer = separation;
                else steer = arrive;
                if (enemy_mode) break;              // enemy_seek_vpref (movement.c:1946): no nullify pass, no fall-backs
                // nullify_impass_components (movement.c:1831)
                if (steer.x > 0 && (!lp || (!on_blocked && lb))) steer.x = 0.0f;
                if (steer.x < 0 && (!rp || (!on_blocked && rb))) steer.x = 0.0f;
                if (steer.z > 0 && (!tpth || (!on_blocked && tb))) steer.z = 0.0f;
                if (steer.z < 0 && (!bp || (!on_blocked && bb))) steer.z = 0.0f;
                if ((double)v2_len(steer) > tp.scaled_max_force_d * 0.01) break;
            }
            const v2 accel = v2_scale(steer, 1.0f / 1.0f);
            vpref = v2_truncate(v2_add(velocity, accel), a.speed / hzf);
            if ((cell_mode || form_mode) && has_form) {
                // formation drag caps the speed at 75 % (movement.c:1940, 2017); unassigned members stand still (:3426, :3438)
                if (v2_len(v2{fin.drag[0], fin.drag[1]}) > EPS_F) vpref = v2_truncate(vpref, (float)(((double)a.speed * 0.75) / (double)hzf));
                if (!ready) vpref = {0.0f, 0.0f};
            } else if (cell_mode || form_mode) vpref = {0.0f, 0.0f};
            }
        }

        // ================= find_neighbours (movement.c:2768) =================
        int ndyn = 0, nstat = 0, raw = 0;
        if (MODE == 2) {
            // phase A stored the neighbour lists (ids, in the reference's order); rebuild the cp_ents
            ndyn = (int)prep[w].ndyn; nstat = (int)prep[w].nstat;
            for (int k = lane; k < ndyn + nstat; k += 32) {
                const bool isst = k >= ndyn;
                const pf_record r = rec[isst ? prep[w].stat_id[k - ndyn] : prep[w].dyn_id[k]];
                cp_ent nd;
                nd.pos = {r.px, r.pz}; nd.radius = r.radius;
                nd.vel = isst ? v2{0.0f, 0.0f} : v2{r.vx, r.vz};
                if (isst) s.stat[k - ndyn] = nd; else s.dyn[k] = nd;
            }
        }
        auto classify = [&](bool hit, uint32_t id) -> bool {
            const uint32_t mk = __ballot_sync(FULL, hit);
            if (!mk) return false;
            const int rrank = __popc(mk & ((1u << lane) - 1));
            bool isdyn = false, isstat = false;
            cp_ent nd;
            if (hit && raw + rrank < 512 && id != uid) {
                const pf_record r = rec[id];
                const uint32_t fl = r.state_flags & 0xFFFFFFu;
                const uint32_t st = r.state_flags >> 24;
                if ((fl & PFNAV_FLAG_MOVABLE) && r.radius != 0.0f &&
                    ((ent_flags & PFNAV_FLAG_AIR) == (fl & PFNAV_FLAG_AIR))) {
                    nd.pos = {r.px, r.pz}; nd.vel = {r.vx, r.vz}; nd.radius = r.radius;
                    const bool still = (st == PFNAV_STATE_ARRIVED || st == PFNAV_STATE_WAITING);
                    if (still || v2_len(nd.vel) < 0.3f) { nd.vel = {0.0f, 0.0f}; isstat = true; }
                    else isdyn = true;
                }
            }
            const uint32_t md = __ballot_sync(FULL, isdyn), ms = __ballot_sync(FULL, isstat);
            if (isdyn) { const int k = ndyn + __popc(md & ((1u << lane) - 1)); if (k < PFNAV_MAX_NEIGHBOURS) { s.dyn[k] = nd; if (MODE == 1) prep[w].dyn_id[k] = id; } }
            if (isstat) { const int k = nstat + __popc(ms & ((1u << lane) - 1)); if (k < PFNAV_MAX_NEIGHBOURS) { s.stat[k] = nd; if (MODE == 1) prep[w].stat_id[k] = id; } }
            ndyn = min(ndyn + __popc(md), PFNAV_MAX_NEIGHBOURS);
            nstat = min(nstat + __popc(ms), PFNAV_MAX_NEIGHBOURS);
            raw += __popc(mk);
            return raw >= 512;
        };
        if (MODE == 2) {
            // nothing to gather
        } else if (!filter_garr) {
            grid_query(g, pos.x, pos.z, 10.0f, lane, classify);
        } else {
            // garrisoned entities exist: G_Pos_EntsInCircleFrom (position.c:379) first takes the raw hits (<= 512),
            // then filter_garrisoned swap-removes them, which reorders the survivors -> materialise the list
            uint32_t *lst = nb_scratch + (size_t)(blockIdx.x * VEL_WARPS_PER_CTA + warp) * 512;
            int cnt = 0;
            grid_query(g, pos.x, pos.z, 10.0f, lane, [&](bool hit, uint32_t id) -> bool {
                const uint32_t mk = __ballot_sync(FULL, hit);
                const int rank = __popc(mk & ((1u << lane) - 1));
                if (hit && cnt + rank < 512) lst[cnt + rank] = id;
                cnt += __popc(mk);
                return cnt >= 512;
            });
            cnt = min(cnt, 512);
            __syncwarp();
            cnt = filter_garrisoned(lst, cnt, rec, lane);
            for (int k0 = 0; k0 < cnt; k0 += 32) {
                const int k = k0 + (int)lane;
                if (classify(k < cnt, k < cnt ? lst[k] : 0u)) break;
            }
        }
        __syncwarp();

        // ================= G_ClearPath_NewVelocity (clearpath.c:694) =================
        const cp_ent self = {{a.prev_pos[0], a.prev_pos[1]}, velocity, a.radius};
        if (MODE == 1) {
            const int nx = clearpath_collect(s, self, ndyn, nstat, lane, prep[w].xp);
            if (lane == 0) {
                prep[w].ndyn = (uint32_t)ndyn; prep[w].nstat = (uint32_t)nstat;
                prep[w].nx = nx > PF_PREP_XP_CAP ? PF_PREP_OVERFLOW : (uint32_t)nx;
            }
            __syncwarp();
            continue;
        }
        v2 new_vel = {0.0f, 0.0f};
        {
            // first solve; when it finds no admissible velocity the drop-furthest retry loop (clearpath.c:702-713) is
            // resolved in one pass over the same ray table (clearpath_retry)
            v2 r;
            int n_rays = 0;
            bool found, exact = true;
            if (MODE == 2 && prep[w].nx != PF_PREP_OVERFLOW)
                found = clearpath_finish(s, self, vpref, ndyn, nstat, lane, prep[w].xp, (int)prep[w].nx, r, n_rays);
            else
                found = clearpath_new_velocity(s, self, vpref, ndyn, nstat, lane, r, n_rays);
            if (found) new_vel = r;
            else {
                if (lane == 0) { atomicAdd(tp.stats + 0, 1ull); if (ndyn > 0 && nstat > 0) atomicAdd(tp.stats + 1, 1ull); }
                new_vel = clearpath_retry(s, self, vpref, ndyn, nstat, n_rays, lane, exact);
                if (!exact) {
                    // replay the reference's loop literally (clearpath.c:702-713)
                    new_vel = {0.0f, 0.0f};
                    unsigned long long solves = 0;
                    while (true) {
                        remove_furthest(s, self.pos, ndyn, nstat, lane);
                        if (!(ndyn > 0 && nstat > 0)) break;
                        int nr2;
                        solves++;
                        if (clearpath_new_velocity(s, self, vpref, ndyn, nstat, lane, r, nr2)) { new_vel = r; break; }
                    }
                    if (lane == 0) { atomicAdd(tp.stats + 2, 1ull); atomicAdd(tp.stats + 3, solves); }
                }
            }
        }
        new_vel = v2_truncate(new_vel, a.max_speed / hzf);      // movement.c:3464
        if (lane == 0) {
            vel_out[w] = make_float2(new_vel.x, new_vel.z);
            vpref_out[w] = make_float2(vpref.x, vpref.z);
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------
#define VEL_SMEM_BYTES ((int)(VEL_WARPS_PER_CTA * sizeof(VelSmem)))
int pfnav_agents_init(pfnav_ctx *ctx)
{
    (void)ctx;
    PF_CUDA(cudaFuncSetAttribute(k_agent_velocity<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, VEL_SMEM_BYTES));
    PF_CUDA(cudaFuncSetAttribute(k_agent_velocity<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, VEL_SMEM_BYTES));
    PF_CUDA(cudaFuncSetAttribute(k_agent_velocity<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, VEL_SMEM_BYTES));
    return 0;
}

void pfnav_agents_free(pfnav_ctx *ctx)
{
    cudaFree(ctx->d_agents); cudaFree(ctx->d_records); cudaFree(ctx->d_flocks);
    cudaFree(ctx->d_flock_start); cudaFree(ctx->d_flock_members); cudaFree(ctx->d_cohesion);
    cudaFree(ctx->d_cell_count); cudaFree(ctx->d_cell_start); cudaFree(ctx->d_cell_fill);
    cudaFree(ctx->d_sorted_ix); cudaFree(ctx->d_sorted_iy); cudaFree(ctx->d_sorted_id); cudaFree(ctx->d_sorted_flock);
    cudaFree(ctx->d_coh_fallback); ctx->d_sorted_flock = nullptr; ctx->d_coh_fallback = nullptr; ctx->cap_coh = 0;
    cudaFree(ctx->d_coh_part); ctx->d_coh_part = nullptr; ctx->cap_coh_part = 0;
    cudaFree(ctx->d_cp_stats); ctx->d_cp_stats = nullptr;
    cudaFree(ctx->d_scan_part); ctx->d_scan_part = nullptr;
    cudaFree(ctx->d_work); cudaFree(ctx->d_vel_out); cudaFree(ctx->d_vpref_out); cudaFree(ctx->d_vdes_out);
    cudaFree(ctx->d_movestate); cudaFree(ctx->d_patches); cudaFree(ctx->d_arrival); cudaFree(ctx->d_nb_scratch);
    cudaFree(ctx->d_member_pos); cudaFree(ctx->d_prep); cudaFree(ctx->d_flock_of); cudaFree(ctx->d_facts);
    cudaFree(ctx->d_formation); cudaFree(ctx->d_ms_ext); cudaFree(ctx->d_enter); cudaFree(ctx->d_ttiles);
    ctx->d_flock_of = nullptr; ctx->d_facts = nullptr; ctx->d_formation = nullptr; ctx->d_ms_ext = nullptr;
    ctx->d_enter = nullptr; ctx->d_ttiles = nullptr; ctx->cap_formation = ctx->cap_ms_ext = ctx->cap_enter = ctx->cap_ttiles = 0;
    if (ctx->update_done) cudaEventDestroy(ctx->update_done);
    cudaFree(ctx->d_los_out); cudaFree(ctx->d_work_count); cudaFree(ctx->d_scan_tmp);
    ctx->d_agents = nullptr; ctx->d_records = nullptr; ctx->d_flocks = nullptr; ctx->d_flock_start = nullptr;
    ctx->d_flock_members = nullptr; ctx->d_cohesion = nullptr; ctx->d_cell_count = nullptr; ctx->d_cell_start = nullptr;
    ctx->d_cell_fill = nullptr; ctx->d_sorted_ix = nullptr; ctx->d_sorted_iy = nullptr; ctx->d_sorted_id = nullptr;
    ctx->d_work = nullptr; ctx->d_vel_out = nullptr; ctx->d_vpref_out = nullptr; ctx->d_vdes_out = nullptr;
    ctx->d_los_out = nullptr; ctx->d_work_count = nullptr; ctx->d_scan_tmp = nullptr;
    ctx->cap_agents = ctx->cap_flocks = ctx->cap_cells = ctx->cap_work = 0;
}

// ---- field pool ----
extern "C" int pfnav_pool_create(pfnav_ctx *ctx, int ndests, int max_fields)
{
    PF_ARG(ctx && ctx->d_cost, "map not created");
    PF_NEED_DEVICE(ctx);
    PF_ARG(ndests > 0 && max_fields > 0, "ndests/max_fields");
    PF_CUDA(cudaSetDevice(ctx->device));
    PF_CUDA(cudaDeviceSynchronize());
    cudaFree(ctx->d_pool_slot); cudaFree(ctx->d_pool_flow); cudaFree(ctx->d_pool_los); cudaFree(ctx->d_pool_touch);
    ctx->d_pool_slot = nullptr; ctx->d_pool_flow = nullptr; ctx->d_pool_los = nullptr; ctx->d_pool_touch = nullptr;
    ctx->los_inflight = false;
    const size_t nslots = (size_t)ndests * ctx->chunk_w * ctx->chunk_h;
    PF_CUDA(cudaMalloc(&ctx->d_pool_slot, nslots * sizeof(int32_t)));
    PF_CUDA(cudaMalloc(&ctx->d_pool_flow, (size_t)max_fields * 4096));
    // LOS region carries a trailing `has` byte per slot
    PF_CUDA(cudaMalloc(&ctx->d_pool_los, (size_t)max_fields * 4096 + max_fields));
    PF_CUDA(cudaMemset(ctx->d_pool_slot, 0xFF, nslots * sizeof(int32_t)));
    PF_CUDA(cudaMemset(ctx->d_pool_los + (size_t)max_fields * 4096, 0, max_fields));
    PF_CUDA(cudaMalloc(&ctx->d_pool_touch, (size_t)max_fields * 4));
    PF_CUDA(cudaMemset(ctx->d_pool_touch, 0, (size_t)max_fields * 4));
    ctx->h_slot_touch.assign(max_fields, 0); ctx->h_slot_owner.assign(max_fields, -1); ctx->pool_free.clear();
    ctx->h_pool_slot.assign(nslots, -1);
    ctx->h_pool_has.assign(max_fields, 0);
    ctx->h_pool_req.assign(max_fields, pfnav_field_req{});
    ctx->h_pool_ffid.assign(nslots, 0);
    ctx->pool_ndests = ndests; ctx->pool_max = max_fields; ctx->pool_used = 0;
    ctx->goal_batch.valid = false;
    return PFNAV_OK;
}

extern "C" int pfnav_pool_clear(pfnav_ctx *ctx)
{
    PF_ARG(ctx && ctx->d_pool_slot, "pool not created");
    PF_CUDA(cudaSetDevice(ctx->device));
    PF_CUDA(pf_fields_sync(ctx));
    PF_CUDA(cudaMemset(ctx->d_pool_slot, 0xFF, ctx->h_pool_slot.size() * sizeof(int32_t)));
    PF_CUDA(cudaMemset(ctx->d_pool_los + (size_t)ctx->pool_max * 4096, 0, ctx->pool_max));
    std::fill(ctx->h_pool_slot.begin(), ctx->h_pool_slot.end(), -1);
    std::fill(ctx->h_pool_has.begin(), ctx->h_pool_has.end(), 0);
    std::fill(ctx->h_pool_ffid.begin(), ctx->h_pool_ffid.end(), 0);
    PF_CUDA(cudaMemset(ctx->d_pool_touch, 0, (size_t)ctx->pool_max * 4));
    std::fill(ctx->h_slot_touch.begin(), ctx->h_slot_touch.end(), 0);
    std::fill(ctx->h_slot_owner.begin(), ctx->h_slot_owner.end(), -1);
    ctx->pool_free.clear();
    ctx->goal_batch.valid = false;
    ctx->pool_used = 0;
    return PFNAV_OK;
}

extern "C" int pfnav_pool_put(pfnav_ctx *ctx, int dest, int chunk_r, int chunk_c, const uint8_t *flow_field,
                              const uint8_t *los_field)
{
    PF_ARG(ctx && ctx->d_pool_slot, "pool not created");
    PF_ARG(dest >= 0 && dest < ctx->pool_ndests, "dest");
    PF_ARG(chunk_r >= 0 && chunk_r < ctx->chunk_h && chunk_c >= 0 && chunk_c < ctx->chunk_w, "chunk");
    PF_CUDA(cudaSetDevice(ctx->device));
    PF_CUDA(pf_fields_sync(ctx));
    const size_t si = (size_t)dest * ctx->chunk_w * ctx->chunk_h + chunk_r * ctx->chunk_w + chunk_c;
    int32_t slot = ctx->h_pool_slot[si];
    uint8_t *d_has = ctx->d_pool_los + (size_t)ctx->pool_max * 4096;
    if (slot < 0) {
        bool evicted = false;
        int rc = pf_pool_reserve(ctx, &si, 1, &slot, &evicted);
        if (rc) return rc;
        if (evicted) {
            PF_CUDA(cudaMemcpy(ctx->d_pool_slot, ctx->h_pool_slot.data(), ctx->h_pool_slot.size() * 4, cudaMemcpyHostToDevice));
            PF_CUDA(cudaMemcpy(d_has, ctx->h_pool_has.data(), ctx->pool_max, cudaMemcpyHostToDevice));
        } else {
            PF_CUDA(cudaMemcpy(ctx->d_pool_slot + si, &slot, sizeof(int32_t), cudaMemcpyHostToDevice));
        }
    }
    uint8_t has = ctx->h_pool_has[slot];
    if (flow_field) { PF_CUDA(cudaMemcpy(ctx->d_pool_flow + (size_t)slot * 4096, flow_field, 4096, cudaMemcpyHostToDevice)); has |= 1; }
    if (los_field) { PF_CUDA(cudaMemcpy(ctx->d_pool_los + (size_t)slot * 4096, los_field, 4096, cudaMemcpyHostToDevice)); has |= 2; }
    ctx->h_pool_has[slot] = has;
    PF_CUDA(cudaMemcpy(d_has + slot, &has, 1, cudaMemcpyHostToDevice));
    ctx->goal_batch.valid = false;
    return PFNAV_OK;
}

// ---- agents ----
template <typename T>
static int ensure(T *&p, size_t &cap, size_t need)
{
    if (cap >= need && p) return 0;
    cudaFree(p);
    p = nullptr;
    PF_CUDA(cudaMalloc(&p, std::max<size_t>(need, 1) * sizeof(T)));
    return 0;
}

static GridView grid_of(const pfnav_ctx *ctx)
{
    GridView g;
    g.cell_start = ctx->d_cell_start; g.cell_count = ctx->d_cell_count;
    g.ix = ctx->d_sorted_ix; g.iy = ctx->d_sorted_iy; g.id = ctx->d_sorted_id;
    g.grid_w = ctx->grid_w; g.grid_h = ctx->grid_h; g.origin_x = ctx->origin_x; g.origin_y = ctx->origin_y;
    return g;
}

static int build_index(pfnav_ctx *ctx, cudaStream_t st)
{
    const int n = (int)ctx->n_agents;
    const int ncells = ctx->grid_w * ctx->grid_h;
    pf_prof_scope prof(ctx, st, PF_PROF_INDEX);
    PF_CUDA(cudaMemsetAsync(ctx->d_cell_count, 0, (size_t)ncells * 4, st));
    PF_CUDA(cudaMemsetAsync(ctx->d_cell_fill, 0, (size_t)ncells * 4, st));
    if (n > 0) {
        const GridView g = grid_of(ctx);
        k_cell_count<<<(n + 255) / 256, 256, 0, st>>>(ctx->d_records, n, g, ctx->d_cell_count);
        const int nsb = (ncells + 4095) / 4096;
        if (nsb > 1 && nsb <= 1024) {
            if (!ctx->d_scan_part) PF_CUDA(cudaMalloc(&ctx->d_scan_part, 1024 * sizeof(uint32_t)));
            k_cell_scan_sum<<<nsb, 1024, 0, st>>>(ctx->d_cell_count, (uint32_t *)ctx->d_scan_part, ncells);
            k_cell_scan_parts<<<1, 1024, 0, st>>>((uint32_t *)ctx->d_scan_part, nsb, ctx->d_cell_start, ncells);
            k_cell_scan_apply<<<nsb, 1024, 0, st>>>(ctx->d_cell_count, (const uint32_t *)ctx->d_scan_part, ctx->d_cell_start, ncells);
            ctx->launches += 2;
        } else
            k_cell_scan<<<1, 1024, 0, st>>>(ctx->d_cell_count, ctx->d_cell_start, ncells);
        k_cell_scatter<<<(n + 255) / 256, 256, 0, st>>>(ctx->d_records, n, g, ctx->d_cell_start, ctx->d_cell_fill,
                                                       ctx->d_sorted_id);
        k_cell_sort<<<(ncells + 127) / 128, 128, 0, st>>>(ctx->d_records, g, ctx->d_cell_start, ctx->d_sorted_id,
                                                         ctx->d_sorted_ix, ctx->d_sorted_iy, ncells, ctx->d_flock_of, ctx->d_sorted_flock);
        ctx->launches += 4;
    } else {
        PF_CUDA(cudaMemsetAsync(ctx->d_cell_start, 0, (size_t)(ncells + 1) * 4, st));
    }
    PF_CUDA(cudaGetLastError());
    return 0;
}

// records of the uploaded range + the flock id column (neighbour kernels read other agents' flock through it)
__global__ void k_make_records(const pfnav_agent *__restrict__ agents, pf_record *__restrict__ rec, int32_t *__restrict__ flock_of,
                               int lo, int hi)
{
    const int i = lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hi) return;
    const pfnav_agent a = agents[i];
    pf_record r;
    r.px = a.pos[0]; r.pz = a.pos[1]; r.vx = a.velocity[0]; r.vz = a.velocity[1];
    r.radius = a.radius;
    r.state_flags = (a.state << 24) | (a.flags & 0xFFFFFFu);
    rec[i] = r;
    flock_of[i] = a.flock;
}

// population-wide facts the tick needs from records it did not upload itself (other ranks' shards): the largest
// selection radius (adjacency pre-selection of k_entity_update) and whether any entity is garrisoned
__global__ void k_population_facts(const pf_record *__restrict__ rec, int n, uint32_t *__restrict__ out)
{
    uint32_t mx = 0, garr = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const pf_record r = rec[i];
        mx = max(mx, __float_as_uint(fmaxf(r.radius, 0.0f)));          // non-negative floats order like their bit patterns
        garr |= (r.state_flags & PFNAV_FLAG_GARRISONED) ? 1u : 0u;
    }
    for (int off = 16; off > 0; off >>= 1) { mx = max(mx, __shfl_xor_sync(FULL, mx, off)); garr |= __shfl_xor_sync(FULL, garr, off); }
    if ((threadIdx.x & 31) == 0) { atomicMax(&out[0], mx); if (garr) atomicOr(&out[1], 1u); }
}

int pfnav_mgpu_allgather(pfnav_ctx *ctx, void *d_buf, size_t elem_bytes, cudaStream_t st);      // pfnav_mgpu.cu

// flock member lists (ascending uid: the iteration order our cohesion sum is defined over) from the flock id column
static int rebuild_flock_members(pfnav_ctx *ctx, cudaStream_t st)
{
    const size_t n = ctx->n_agents, nflocks = ctx->n_flocks;
    std::vector<int32_t> fo(n ? n : 1);
    PF_CUDA(cudaMemcpyAsync(fo.data(), ctx->d_flock_of, n * 4, cudaMemcpyDeviceToHost, st));
    PF_CUDA(cudaStreamSynchronize(st));
    std::vector<uint32_t> fstart(nflocks + 1, 0), members(n ? n : 1);
    for (size_t i = 0; i < n; i++) {
        PF_ARG(fo[i] < (int)nflocks, "agent flock index out of range");
        if (fo[i] >= 0) fstart[fo[i] + 1]++;
    }
    for (size_t f = 0; f < nflocks; f++) fstart[f + 1] += fstart[f];
    {
        std::vector<uint32_t> cur(fstart.begin(), fstart.end() - 1);
        for (size_t i = 0; i < n; i++)
            if (fo[i] >= 0) members[cur[fo[i]]++] = (uint32_t)i;
    }
    PF_CUDA(cudaMemcpyAsync(ctx->d_flock_start, fstart.data(), (nflocks + 1) * 4, cudaMemcpyHostToDevice, st));
    if (n) PF_CUDA(cudaMemcpyAsync(ctx->d_flock_members, members.data(), n * 4, cudaMemcpyHostToDevice, st));
    PF_CUDA(cudaStreamSynchronize(st));      // host vectors go out of scope
    ctx->h_flock_start = fstart;
    return 0;
}

// move_copy_gamestate (movement.c:3607) for the index range [lo, hi) of a population of n_total entities
// (uid == index into the whole population). Single GPU: lo = 0, hi = n_total. Multi-GPU: every rank uploads its own
// range (pfnav_mgpu_shard_range) and the 24-byte neighbour records + flock ids of the other ranges arrive through
// the all-gather.
static int agents_upload_impl(pfnav_ctx *ctx, const pfnav_agent *agents, size_t lo, size_t hi, size_t n,
                              const pfnav_flock *flocks, size_t nflocks, int hz, uint32_t flags)
{
    PF_ARG(ctx && ctx->d_cost, "map not created");
    PF_NEED_DEVICE(ctx);
    PF_ARG(lo <= hi && hi <= n, "shard range");
    PF_ARG(hi == lo || agents, "agents");
    PF_ARG(nflocks == 0 || flocks, "flocks");
    PF_ARG(hz == 20 || hz == 10 || hz == 5 || hz == 1, "hz must be 20, 10, 5 or 1 (movement.c:2210)");
    PF_ARG(n < (1u << 31), "n");
    const bool sharded = ctx->mgpu != nullptr;
    PF_ARG(sharded || (lo == 0 && hi == n), "a partial upload needs pfnav_mgpu_init / pfnav_group_create first");
    PF_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->tick_stream;
    const bool same = (flags & PFNAV_UPLOAD_SAME_FLOCKS) && ctx->n_agents == n && ctx->n_flocks == nflocks &&
                      ctx->shard_lo == lo && ctx->shard_hi == hi && ctx->hz == hz;
    ctx->hz = hz;
    int rc;
    if (n > ctx->cap_agents) {
        size_t cap = 0;
        if ((rc = ensure(ctx->d_agents, cap, n))) return rc; cap = 0;
        if ((rc = ensure(ctx->d_records, cap, n))) return rc; cap = 0;
        if ((rc = ensure(ctx->d_flock_members, cap, n))) return rc; cap = 0;
        if ((rc = ensure(ctx->d_flock_of, cap, n))) return rc; cap = 0;
        if ((rc = ensure(ctx->d_sorted_ix, cap, n))) return rc; cap = 0;
        if ((rc = ensure(ctx->d_sorted_iy, cap, n))) return rc; cap = 0;
        if ((rc = ensure(ctx->d_sorted_flock, cap, n))) return rc; cap = 0;
        if ((rc = ensure(ctx->d_sorted_id, cap, n))) return rc;
        ctx->cap_agents = n;
    }
    if (nflocks + 1 > ctx->cap_flocks) {
        size_t cap = 0;
        if ((rc = ensure(ctx->d_flocks, cap, nflocks + 1))) return rc; cap = 0;
        if ((rc = ensure(ctx->d_flock_start, cap, nflocks + 2))) return rc;
        ctx->cap_flocks = nflocks + 1;
    }
    if (!ctx->d_facts) PF_CUDA(cudaMalloc(&ctx->d_facts, 8));
    ctx->n_agents = n; ctx->n_flocks = nflocks; ctx->shard_lo = lo; ctx->shard_hi = hi;
    if (!same) {
        // what the host needs to know about the entities it will run work items for (its own range)
        ctx->flock_layer_used.assign(nflocks * PFNAV_NAV_LAYER_MAX, 0);
        ctx->has_unsupported_state = false;
        for (size_t i = 0; i < hi - lo; i++) {
            PF_ARG(agents[i].flock < (int)nflocks, "agent flock index out of range");
            // Entity_NavLayerWithRadius (entity.c:554): the layer this agent's tile probes read
            const uint32_t f = agents[i].flags; const float r = agents[i].radius;
            const int base = (f & PFNAV_FLAG_WATER) ? 4 : (f & PFNAV_FLAG_AIR) ? 8 : 0;
            const int layer = base + (r >= 15.0f ? 3 : r >= 10.0f ? 2 : r >= 5.0f ? 1 : 0);
            PF_ARG(layer < ctx->nlayers, "agent needs a navigation layer that was not created (radius/flags)");
            if (agents[i].flock >= 0) ctx->flock_layer_used[(size_t)agents[i].flock * PFNAV_NAV_LAYER_MAX + layer] = 1;
            const uint32_t stt = agents[i].state;
            if (stt == PFNAV_STATE_MOVING_IN_FORMATION || stt == PFNAV_STATE_SURROUND_ENTITY ||
                stt == PFNAV_STATE_ENTER_ENTITY_RANGE || stt == PFNAV_STATE_TURNING || stt == PFNAV_STATE_ARRIVING_TO_CELL)
                ctx->has_unsupported_state = true;
        }
        ctx->h_flocks.assign(flocks, flocks + nflocks);
        ctx->arrival_valid = false;
        // optional per-entity inputs of the previous population do not carry over
        cudaFree(ctx->d_formation); ctx->d_formation = nullptr; ctx->cap_formation = 0;
        cudaFree(ctx->d_ms_ext); ctx->d_ms_ext = nullptr; ctx->cap_ms_ext = 0;
    }
    if (hi > lo) PF_CUDA(cudaMemcpyAsync(ctx->d_agents + lo, agents, (hi - lo) * sizeof(pfnav_agent), cudaMemcpyHostToDevice, st));
    if (nflocks && !same) PF_CUDA(cudaMemcpyAsync(ctx->d_flocks, flocks, nflocks * sizeof(pfnav_flock), cudaMemcpyHostToDevice, st));
    // spatial index geometry: G_Pos_Init (position.c:264) + bg_init (bitmap_grid.h:959)
    {
        const float W = (float)(ctx->chunk_w * 256), H = (float)(ctx->chunk_h * 256);
        const float cx = ctx->map_x - W / 2.0f, cz = ctx->map_z + H / 2.0f;
        const float xmin = cx - W / 2.0f, xmax = cx + W / 2.0f, zmin = cz - H / 2.0f, zmax = cz + H / 2.0f;
        ctx->origin_x = (int32_t)lrintf(xmin * 256.0f);
        ctx->origin_y = (int32_t)lrintf(zmin * 256.0f);
        const int32_t span_x = (int32_t)lrintf(xmax * 256.0f) - ctx->origin_x;
        const int32_t span_y = (int32_t)lrintf(zmax * 256.0f) - ctx->origin_y;
        ctx->grid_w = std::max(1, (int)(((uint32_t)span_x + 4095u) >> 12));
        ctx->grid_h = std::max(1, (int)(((uint32_t)span_y + 4095u) >> 12));
        const size_t ncells = (size_t)ctx->grid_w * ctx->grid_h;
        if (ncells + 1 > ctx->cap_cells) {
            size_t cap = 0;
            if ((rc = ensure(ctx->d_cell_count, cap, ncells + 1))) return rc; cap = 0;
            if ((rc = ensure(ctx->d_cell_start, cap, ncells + 1))) return rc; cap = 0;
            if ((rc = ensure(ctx->d_cell_fill, cap, ncells + 1))) return rc;
            ctx->cap_cells = ncells + 1;
        }
    }
    if (hi > lo) {
        k_make_records<<<((int)(hi - lo) + 255) / 256, 256, 0, st>>>(ctx->d_agents, ctx->d_records, ctx->d_flock_of, (int)lo, (int)hi);
        ctx->launches++;
    }
    // default work list: none until pfnav_agents_set_work (a same-structure re-upload keeps the current one)
    if (!same) ctx->n_work = 0;
    if (sharded) {
        // the other ranges' records (and, when membership may have changed, flock ids) come from the peers;
        // pfnav_mgpu_gather / pfnav_group_gather follows and finishes the snapshot (index, facts, member lists)
        ctx->members_stale = !same;
        PF_CUDA(cudaStreamSynchronize(st));
        return PFNAV_OK;
    }
    return pfnav_agents_finish_snapshot(ctx, st, !same);
}

// Everything that needs the WHOLE population's records: member lists, population facts, spatial index.
int pfnav_agents_finish_snapshot(pfnav_ctx *ctx, cudaStream_t st, bool members)
{
    int rc;
    const int n = (int)ctx->n_agents;
    if (members && (rc = rebuild_flock_members(ctx, st))) return rc;
    if (members) {
        PF_CUDA(cudaMemsetAsync(ctx->d_facts, 0, 8, st));
        if (n) { k_population_facts<<<std::min((n + 255) / 256, 1024), 256, 0, st>>>(ctx->d_records, n, ctx->d_facts); ctx->launches++; }
        uint32_t facts[2] = {0, 0};
        PF_CUDA(cudaMemcpyAsync(facts, ctx->d_facts, 8, cudaMemcpyDeviceToHost, st));
        PF_CUDA(cudaStreamSynchronize(st));
        memcpy(&ctx->max_radius, &facts[0], 4);
        ctx->any_garrisoned = facts[1] != 0;
    }
    if ((rc = build_index(ctx, st))) return rc;
    PF_CUDA(cudaStreamSynchronize(st));
    return PFNAV_OK;
}

extern "C" int pfnav_agents_upload(pfnav_ctx *ctx, const pfnav_agent *agents, size_t n, const pfnav_flock *flocks,
                                   size_t nflocks, int hz)
{
    PF_ARG(ctx && !ctx->mgpu, "multi-GPU contexts upload their own range: pfnav_agents_upload_shard");
    return agents_upload_impl(ctx, agents, 0, n, n, flocks, nflocks, hz, 0);
}

extern "C" int pfnav_agents_upload_shard(pfnav_ctx *ctx, const pfnav_agent *shard, size_t lo, size_t hi, size_t n_total,
                                         const pfnav_flock *flocks, size_t nflocks, int hz, uint32_t flags)
{
    PF_ARG(ctx, "ctx");
    return agents_upload_impl(ctx, shard, lo, hi, n_total, flocks, nflocks, hz, flags);
}

extern "C" int pfnav_agents_rebuild_index(pfnav_ctx *ctx, void *stream)
{
    PF_ARG(ctx && ctx->d_records, "agents not uploaded");
    PF_CUDA(cudaSetDevice(ctx->device));
    return build_index(ctx, pf_stream(ctx, stream));
}

extern "C" int pfnav_agents_device_ptrs(pfnav_ctx *ctx, void **d_records, void **d_velocities, size_t *n)
{
    PF_ARG(ctx, "ctx");
    if (d_records) *d_records = ctx->d_records;
    if (d_velocities) *d_velocities = ctx->d_vel_out;
    if (n) *n = ctx->n_agents;
    return PFNAV_OK;
}

extern "C" int pfnav_agents_set_work(pfnav_ctx *ctx, const uint32_t *uids, size_t nwork)
{
    PF_ARG(ctx && ctx->d_agents, "agents not uploaded");
    PF_CUDA(cudaSetDevice(ctx->device));
    std::vector<uint32_t> all;
    if (!uids) {
        // every agent that is not still (ent_still, movement.c:652) -- needs the host copy of state
        const size_t lo = ctx->shard_lo, cnt = ctx->shard_hi - ctx->shard_lo;
        std::vector<pfnav_agent> h(cnt);
        PF_CUDA(cudaDeviceSynchronize());
        if (cnt) PF_CUDA(cudaMemcpy(h.data(), ctx->d_agents + lo, cnt * sizeof(pfnav_agent), cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < cnt; i++)
            if (h[i].state != PFNAV_STATE_ARRIVED && h[i].state != PFNAV_STATE_WAITING) all.push_back((uint32_t)(lo + i));
        uids = all.data();
        nwork = all.size();
    }
    for (size_t i = 0; i < nwork; i++)
        PF_ARG(uids[i] >= ctx->shard_lo && uids[i] < ctx->shard_hi, "work uid outside this context's own entity range");
    if (nwork > ctx->cap_work) {
        size_t cap = 0; int rc;
        if ((rc = ensure(ctx->d_work, cap, nwork))) return rc; cap = 0;
        if ((rc = ensure(ctx->d_vel_out, cap, nwork))) return rc; cap = 0;
        if ((rc = ensure(ctx->d_vpref_out, cap, nwork))) return rc; cap = 0;
        if ((rc = ensure(ctx->d_vdes_out, cap, nwork))) return rc; cap = 0;
        if ((rc = ensure(ctx->d_los_out, cap, nwork))) return rc;
        ctx->cap_work = nwork;
    }
    if (!ctx->d_work_count) PF_CUDA(cudaMalloc(&ctx->d_work_count, 16));
    if (nwork) PF_CUDA(cudaMemcpy(ctx->d_work, uids, nwork * 4, cudaMemcpyHostToDevice));
    ctx->n_work = nwork;
    return PFNAV_OK;
}

__global__ void k_copy_vdes(const pfnav_agent *__restrict__ agents, const uint32_t *__restrict__ work, int nwork,
                            float2 *__restrict__ vdes, uint8_t *__restrict__ los)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= nwork) return;
    const pfnav_agent &a = agents[work[w]];
    vdes[w] = make_float2(a.vdes[0], a.vdes[1]);
    los[w] = a.has_dest_los ? 1 : 0;
}

extern "C" int pfnav_agents_tick(pfnav_ctx *ctx, uint32_t flags, void *stream)
{
    PF_ARG(ctx && ctx->d_agents, "agents not uploaded");
    if (ctx->n_work == 0) return PFNAV_OK;
    PF_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = pf_stream(ctx, stream);
    const int nwork = (int)ctx->n_work;
    ctx->tick_no++;
    MapView m;
    m.cost = ctx->d_cost; m.blk = ctx->d_blk; m.W64 = ctx->W64; m.H64 = ctx->H64;
    m.chunk_w = ctx->chunk_w; m.chunk_h = ctx->chunk_h; m.map_x = ctx->map_x; m.map_z = ctx->map_z;
    // SCALED_MAX_FORCE = (MAX_FORCE / hz_count(hz) * 20.0): float division, double product (movement.c:93)
    TickParams tp;
    tp.hz = ctx->hz;
    tp.scaled_max_force_d = (double)(0.75f / (float)ctx->hz) * 20.0;
    tp.scaled_max_force = (float)tp.scaled_max_force_d;
    if (!ctx->d_cp_stats) {
        PF_CUDA(cudaMalloc(&ctx->d_cp_stats, 16 * sizeof(unsigned long long)));
        PF_CUDA(cudaMemsetAsync(ctx->d_cp_stats, 0, 16 * sizeof(unsigned long long), st));
    }
    tp.stats = (unsigned long long *)ctx->d_cp_stats;
    tp.queue = (unsigned int *)(tp.stats + 8);
    PF_CUDA(cudaMemsetAsync(tp.queue, 0, 4 * sizeof(unsigned int), st));
    PF_CUDA(cudaMemsetAsync(ctx->d_work_count, 0, 4, st));
    {
    pf_prof_scope prof(ctx, st, PF_PROF_COHESION);
    if (ctx->cap_member_pos < ctx->n_agents) {
        cudaFree(ctx->d_member_pos); ctx->d_member_pos = nullptr; ctx->cap_member_pos = 0;
        PF_CUDA(cudaMalloc(&ctx->d_member_pos, std::max<size_t>(ctx->n_agents, 1) * sizeof(float2)));
        ctx->cap_member_pos = ctx->n_agents;
    }
    if (ctx->cap_coh < ctx->n_agents) {
        cudaFree(ctx->d_cohesion); cudaFree(ctx->d_coh_fallback); ctx->d_cohesion = nullptr; ctx->d_coh_fallback = nullptr; ctx->cap_coh = 0;
        PF_CUDA(cudaMalloc(&ctx->d_cohesion, std::max<size_t>(ctx->n_agents, 1) * sizeof(float2)));
        PF_CUDA(cudaMalloc(&ctx->d_coh_fallback, (std::max<size_t>(ctx->n_agents, 1) + 4) * sizeof(uint32_t)));
        ctx->cap_coh = ctx->n_agents;
    }
    k_gather_flock_pos<<<((int)ctx->n_agents + 255) / 256, 256, 0, st>>>(ctx->d_records, ctx->d_flock_members, (int)ctx->n_agents,
                                                                        ctx->d_member_pos);
    // big flocks: windowed pass over the position index + full-list fall-back for entities without company; small
    // populations: the full member list directly (the window would cover the whole flock anyway)
    size_t biggest = 0;
    for (size_t f = 0; f + 1 < ctx->h_flock_start.size(); f++) biggest = std::max<size_t>(biggest, ctx->h_flock_start[f + 1] - ctx->h_flock_start[f]);
    const bool windowed = ctx->cohesion_mode == 1 || (ctx->cohesion_mode == 0 && biggest >= 20000);
    if (windowed) {
        uint32_t *d_nfb = ctx->d_coh_fallback + ctx->n_agents;
        PF_CUDA(cudaMemsetAsync(d_nfb, 0, 4, st));
        const int nblk = ((int)ctx->n_agents + COH_THREADS - 1) / COH_THREADS;     // one 128-entry run each: the hardware scheduler balances them
        // fewer runs than ~8 blocks per SM: split every run's window rows over blockIdx.y
        const int nsplit = std::max(1, std::min(8, (ctx->sm_count * 8 + nblk - 1) / nblk));
        if (nsplit > 1 && ctx->cap_coh_part < (size_t)nsplit * ctx->n_agents) {
            cudaFree(ctx->d_coh_part); ctx->d_coh_part = nullptr; ctx->cap_coh_part = 0;
    cudaFree(ctx->d_cp_stats); ctx->d_cp_stats = nullptr;
    cudaFree(ctx->d_scan_part); ctx->d_scan_part = nullptr;
            PF_CUDA(cudaMalloc(&ctx->d_coh_part, (size_t)8 * ctx->n_agents * sizeof(float4)));
            ctx->cap_coh_part = (size_t)8 * ctx->n_agents;
        }
        k_cohesion_window<<<dim3(nblk, nsplit), COH_THREADS, 0, st>>>(grid_of(ctx), ctx->d_records, ctx->d_sorted_flock, ctx->d_flock_start,
                                                        (int)ctx->n_agents, (uint32_t)ctx->shard_lo, (uint32_t)ctx->shard_hi,
                                                        tp.scaled_max_force, ctx->d_cohesion, ctx->d_coh_fallback, d_nfb,
                                                        (float4 *)ctx->d_coh_part, nsplit);
        if (nsplit > 1) {
            k_cohesion_finish<<<((int)ctx->n_agents + 127) / 128, 128, 0, st>>>(grid_of(ctx), ctx->d_records, ctx->d_sorted_flock,
                ctx->d_flock_start, (int)ctx->n_agents, (uint32_t)ctx->shard_lo, (uint32_t)ctx->shard_hi, tp.scaled_max_force,
                ctx->d_cohesion, ctx->d_coh_fallback, d_nfb, (const float4 *)ctx->d_coh_part, nsplit);
            ctx->launches++;
        }
        // fall-back entities: the grid is sized for the worst case, the kernel reads the count on the device
        const int nown = (int)(ctx->shard_hi - ctx->shard_lo);
        k_cohesion<<<(nown + 127) / 128, 128, 0, st>>>(ctx->d_records, ctx->d_flock_of, ctx->d_flock_start, ctx->d_member_pos,
                                                      ctx->d_coh_fallback, d_nfb, 0, tp.scaled_max_force, ctx->d_cohesion);
        ctx->launches += 2;
    } else {
        k_cohesion<<<(nwork + 127) / 128, 128, 0, st>>>(ctx->d_records, ctx->d_flock_of, ctx->d_flock_start, ctx->d_member_pos,
                                                       ctx->d_work, nullptr, nwork, tp.scaled_max_force, ctx->d_cohesion);
        ctx->launches++;
    }
    }
    const int ctas = std::min((nwork + VEL_WARPS_PER_CTA - 1) / VEL_WARPS_PER_CTA, ctx->sm_count * 8 * 4);
    if (ctx->any_garrisoned && ctx->nb_scratch_warps < (size_t)ctas * VEL_WARPS_PER_CTA) {
        // per-warp raw neighbour lists for the garrisoned-filter path (find_neighbours, 512 ids each)
        cudaFree(ctx->d_nb_scratch); ctx->d_nb_scratch = nullptr; ctx->nb_scratch_warps = 0;
        PF_CUDA(cudaMalloc(&ctx->d_nb_scratch, (size_t)ctas * VEL_WARPS_PER_CTA * 512 * sizeof(uint32_t)));
        ctx->nb_scratch_warps = (size_t)ctas * VEL_WARPS_PER_CTA;
    }
    // Two-phase velocity update while LOS chains are still in flight on the field stream: the part of
    // ClearPath that does not depend on the preferred velocity (neighbours, velocity obstacles, admissible
    // ray intersections) runs now, next to them; only the choice waits for the fields.
    // Policy: splitting costs ~5-30 % extra work (phase B rebuilds the obstacles), so it pays only while the LOS
    // chains outlast a good part of phase A. Both durations are measured on the previous tick.
    if (ctx->vel_timed && cudaEventQuery(ctx->ev_vel1) == cudaSuccess) {
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, ctx->ev_vel0, ctx->ev_vel1) == cudaSuccess) ctx->last_vel_ms = ms;
    }
    bool worth = true;
    if (ctx->last_los_ms > 0.0f && ctx->last_vel_ms > 0.0f) worth = ctx->last_los_ms > 0.5f * ctx->last_vel_ms;
    bool two_phase = ctx->two_phase && (ctx->two_phase_force ||
                                        (worth && ctx->los_inflight && cudaEventQuery(ctx->ev_los) == cudaErrorNotReady));
    cudaGetLastError();
    PF_CUDA(cudaEventRecord(ctx->ev_vel0, st));
    if (two_phase && ctx->cap_prep < (size_t)nwork) {
        cudaFree(ctx->d_prep); ctx->d_prep = nullptr; ctx->cap_prep = 0;
        if (cudaMalloc(&ctx->d_prep, (size_t)nwork * sizeof(pf_prep)) == cudaSuccess) ctx->cap_prep = (size_t)nwork;
        else { cudaGetLastError(); two_phase = false; }         // no room: fall back to the single pass
    }
    if (two_phase) {
        pf_prof_scope prof(ctx, st, PF_PROF_VELOCITY);
        // phase A has the whole LOS phase to hide in: a smaller persistent grid leaves the schedulers to the
        // latency-bound LOS threads it shares the SMs with
        const int ctas_a = ctx->phase_a_ctas_per_sm > 0 ? std::min(ctas, ctx->sm_count * ctx->phase_a_ctas_per_sm) : ctas;
        k_agent_velocity<1><<<ctas_a, VEL_WARPS_PER_CTA * 32, VEL_SMEM_BYTES, st>>>(m, grid_of(ctx), tp, ctx->d_agents, ctx->d_records, ctx->d_flocks,
                                                                    ctx->d_work, nwork, ctx->d_vdes_out, ctx->d_los_out,
                                                                    ctx->d_cohesion, ctx->d_vel_out, ctx->d_vpref_out,
                                                                    ctx->any_garrisoned ? 1 : 0, ctx->d_nb_scratch, (pf_prep *)ctx->d_prep, ctx->d_formation);
        ctx->launches++;
    }
    // everything above is independent of the flow/LOS fields; the LOS chains forked by
    // pfnav_pool_request_goals have been running alongside it
    PF_CUDA(pf_fields_join(ctx, st));
    {
    pf_prof_scope prof(ctx, st, PF_PROF_VDES);
    if (flags & PFNAV_TICK_VDES_FROM_POOL) {
        PF_ARG(ctx->d_pool_slot, "PFNAV_TICK_VDES_FROM_POOL needs a field pool");
        PoolView pv;
        pv.slot = ctx->d_pool_slot; pv.flow = ctx->d_pool_flow; pv.los = ctx->d_pool_los;
        pv.has = ctx->d_pool_los + (size_t)ctx->pool_max * 4096; pv.ndests = ctx->pool_ndests;
        pv.touch = ctx->d_pool_touch; pv.tick_no = ctx->tick_no;
        k_desired_velocity<<<(nwork + 127) / 128, 128, 0, st>>>(m, pv, ctx->d_agents, ctx->d_flocks, ctx->d_work, nwork,
                                                               ctx->d_vdes_out, ctx->d_los_out, ctx->d_work_count,
                                                               ctx->d_ms_ext, ctx->d_formation);
    } else {
        k_copy_vdes<<<(nwork + 255) / 256, 256, 0, st>>>(ctx->d_agents, ctx->d_work, nwork, ctx->d_vdes_out, ctx->d_los_out);
    }
    }
    pf_prof_scope prof(ctx, st, PF_PROF_VELOCITY);
    if (two_phase)
        k_agent_velocity<2><<<ctas, VEL_WARPS_PER_CTA * 32, VEL_SMEM_BYTES, st>>>(m, grid_of(ctx), tp, ctx->d_agents, ctx->d_records, ctx->d_flocks,
                                                                    ctx->d_work, nwork, ctx->d_vdes_out, ctx->d_los_out,
                                                                    ctx->d_cohesion, ctx->d_vel_out, ctx->d_vpref_out,
                                                                    ctx->any_garrisoned ? 1 : 0, ctx->d_nb_scratch, (pf_prep *)ctx->d_prep, ctx->d_formation);
    else
        k_agent_velocity<0><<<ctas, VEL_WARPS_PER_CTA * 32, VEL_SMEM_BYTES, st>>>(m, grid_of(ctx), tp, ctx->d_agents, ctx->d_records, ctx->d_flocks,
                                                                    ctx->d_work, nwork, ctx->d_vdes_out, ctx->d_los_out,
                                                                    ctx->d_cohesion, ctx->d_vel_out, ctx->d_vpref_out,
                                                                    ctx->any_garrisoned ? 1 : 0, ctx->d_nb_scratch, nullptr, ctx->d_formation);
    ctx->launches += 3;
    PF_CUDA(cudaGetLastError());
    PF_CUDA(cudaEventRecord(ctx->ev_vel1, st));
    ctx->vel_timed = true;
    prof.~pf_prof_scope(); prof.a = nullptr;
    PF_CUDA(cudaEventRecord(ctx->tick_done, st));
    return PFNAV_OK;
}

extern "C" int pfnav_agents_read_velocities(pfnav_ctx *ctx, float *out_xz, size_t maxout)
{
    PF_ARG(ctx && out_xz, "null");
    PF_CUDA(cudaSetDevice(ctx->device));
    PF_CUDA(cudaEventSynchronize(ctx->tick_done));
    const size_t n = std::min(maxout, ctx->n_work);
    if (n) PF_CUDA(cudaMemcpy(out_xz, ctx->d_vel_out, n * 8, cudaMemcpyDeviceToHost));
    return PFNAV_OK;
}

extern "C" int pfnav_agents_read_debug(pfnav_ctx *ctx, float *out_vpref_xz, float *out_vdes_xz, uint8_t *out_has_los,
                                       size_t maxout)
{
    PF_ARG(ctx, "null");
    PF_CUDA(cudaSetDevice(ctx->device));
    PF_CUDA(cudaEventSynchronize(ctx->tick_done));
    const size_t n = std::min(maxout, ctx->n_work);
    if (n && out_vpref_xz) PF_CUDA(cudaMemcpy(out_vpref_xz, ctx->d_vpref_out, n * 8, cudaMemcpyDeviceToHost));
    if (n && out_vdes_xz) PF_CUDA(cudaMemcpy(out_vdes_xz, ctx->d_vdes_out, n * 8, cudaMemcpyDeviceToHost));
    if (n && out_has_los) PF_CUDA(cudaMemcpy(out_has_los, ctx->d_los_out, n, cudaMemcpyDeviceToHost));
    return PFNAV_OK;
}

extern "C" int pfnav_ents_in_circle(pfnav_ctx *ctx, float x, float z, float range, uint32_t *out, int maxout, int *out_n)
{
    PF_ARG(ctx && ctx->d_records && out && out_n && maxout > 0, "args");
    PF_CUDA(cudaSetDevice(ctx->device));
    if (range < 0.0f) { *out_n = 0; return PFNAV_OK; }
    uint32_t *d_out = nullptr; int *d_n = nullptr;
    PF_CUDA(cudaMalloc(&d_out, (size_t)maxout * 4));
    PF_CUDA(cudaMalloc(&d_n, 4));
    k_ents_in_circle<<<1, 32, 0, ctx->tick_stream>>>(grid_of(ctx), x, z, range, d_out, maxout, d_n, ctx->d_records,
                                                     ctx->any_garrisoned ? 1 : 0);
    ctx->launches++;
    cudaError_t e = cudaStreamSynchronize(ctx->tick_stream);
    if (e == cudaSuccess) e = cudaMemcpy(out_n, d_n, 4, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && *out_n > 0) e = cudaMemcpy(out, d_out, (size_t)*out_n * 4, cudaMemcpyDeviceToHost);
    cudaFree(d_out); cudaFree(d_n);
    if (e != cudaSuccess) { pfnav_set_error("pfnav_ents_in_circle: %s", cudaGetErrorString(e)); return PFNAV_ERR_CUDA; }
    return PFNAV_OK;
}


// ------------------------------------------------------------------------------------------
// State update: entity_compute_update (movement.c:2303-2650) + the movestate part of
// entity_apply_update (movement.c:2693-2757). One thread per work item.
// The quaternion helpers keep the reference's mix of float and double arithmetic (pf_math.c).
// ------------------------------------------------------------------------------------------
struct quat { float x, y, z, w; };
struct pf_arrival_dev { int32_t nearest_ok; float nearest[2]; int32_t mc_n; int32_t mc_off; int32_t _pad[3]; };

#define PF_PI_D 3.14159265358979323846

// first column of PFM_Mat4x4_RotFromQuat (pf_math.c:324) applied to (1,0,0,1) by PFM_Quat_PitchDiff (:677)
__device__ __forceinline__ void quat_front(const quat &q, float &dx, float &dz)
{
    dx = (float)(1 - 2 * ((double)q.y * (double)q.y) - 2 * ((double)q.z * (double)q.z));   // pow(y,2) is exact y*y in double
    dz = 2 * q.x * q.z + 2 * q.w * q.y;
    // dir_homo.w == 1; the zero products of PFM_Mat4x4_Mult4x1 add nothing for finite inputs
}

// PFM_Quat_PitchDiff (pf_math.c:677-704)
__device__ __forceinline__ float quat_pitch_diff(const quat &a, const quat &b)
{
    float ax, az, bx, bz;
    quat_front(a, ax, az);
    quat_front(b, bx, bz);
    const float dot = ax * bx + az * bz;
    const float det = ax * bz - az * bx;
    return (float)atan2((double)det, (double)dot);
}

// dir_quat_from_velocity (movement.c:1411)
__device__ __forceinline__ quat dir_quat_from_velocity(v2 v)
{
    const float angle_rad = (float)(atan2((double)v.z, (double)v.x) - PF_PI_D / 2.0f);
    quat q;
    q.x = 0.0f; q.z = 0.0f;
    q.y = (float)(1.0f * sin((double)(angle_rad / 2.0f)));
    q.w = (float)cos((double)(angle_rad / 2.0f));
    return q;
}

// turn_toward (movement.c:2249): PFM_Mat4x4_MakeRotY -> PFM_Quat_FromRotMat -> MultQuat -> Normal
__device__ quat turn_toward(const quat &cur, const quat &target, float max_deg)
{
    float angle_deg = (float)((double)quat_pitch_diff(cur, target) * (180.0f / PF_PI_D));
    if (180.0f - fabs((double)angle_deg) < 1.0f) angle_deg = 180.0f;
    const double mn = ((double)max_deg < fabs((double)angle_deg)) ? (double)max_deg : fabs((double)angle_deg);
    const int sg = (angle_deg > 0) - (angle_deg < 0);
    const float turn_deg = (float)(mn * -sg);
    const float radians = (float)((double)turn_deg * (PF_PI_D / 180.0f));
    const float c = (float)cos((double)radians), sn = (float)sin((double)radians);
    // rotation matrix about Y: cols[0][0] = c, cols[0][2] = -s, cols[2][0] = s, cols[2][2] = c, cols[1][1] = 1
    const float m00 = c, m02 = -sn, m20 = sn, m22 = c, m11 = 1.0f;
    quat rot;
    const float tr = m00 + m11 + m22;
    if (tr > 0) {
        const float S = (float)(sqrt((double)tr + 1.0) * 2);
        rot.w = (float)(0.25 * (double)S);
        rot.x = (0.0f - 0.0f) / S;
        rot.y = (m02 - m20) / S;
        rot.z = (0.0f - 0.0f) / S;
    } else if ((m00 > m11) && (m00 > m22)) {
        const float S = (float)(sqrt(1.0 + (double)m00 - (double)m11 - (double)m22) * 2);
        rot.w = (0.0f - 0.0f) / S; rot.x = (float)(0.25 * (double)S); rot.y = (0.0f + 0.0f) / S; rot.z = (m02 + m20) / S;
    } else if (m11 > m22) {
        const float S = (float)(sqrt(1.0 + (double)m11 - (double)m00 - (double)m22) * 2);
        rot.w = (m02 - m20) / S; rot.x = (0.0f + 0.0f) / S; rot.y = (float)(0.25 * (double)S); rot.z = (0.0f + 0.0f) / S;
    } else {
        const float S = (float)(sqrt(1.0 + (double)m22 - (double)m00 - (double)m11) * 2);
        rot.w = (0.0f - 0.0f) / S; rot.x = (m02 + m20) / S; rot.y = (0.0f + 0.0f) / S; rot.z = (float)(0.25 * (double)S);
    }
    // PFM_Quat_MultQuat(&rot, &cur) (pf_math.c:639)
    quat f;
    f.x = ( rot.x * cur.w) + (rot.y * cur.z) - (rot.z * cur.y) + (rot.w * cur.x);
    f.y = (-rot.x * cur.z) + (rot.y * cur.w) + (rot.z * cur.x) + (rot.w * cur.y);
    f.z = ( rot.x * cur.y) - (rot.y * cur.x) + (rot.z * cur.w) + (rot.w * cur.z);
    f.w = (-rot.x * cur.x) - (rot.y * cur.y) - (rot.z * cur.z) + (rot.w * cur.w);
    const float len = (float)sqrt((double)(f.x * f.x + f.y * f.y + f.z * f.z + f.w * f.w));
    f.x = f.x / len; f.y = f.y / len; f.z = f.z / len; f.w = f.w / len;
    return f;
}

// n_tile_blocked (nav.c:235)
__device__ __forceinline__ bool tile_blocked_abs(const MapView &m, int layer, int ar, int ac)
{
    const size_t off = ((size_t)layer * m.H64 + ar) * m.W64 + ac;
    return m.cost[off] == 0xFF || m.blk[off] > 0;
}

struct UpdateParams { int hz; float turn_rate; float adj_query_r; };

__global__ void __launch_bounds__(128)
k_entity_update(MapView m, GridView g, UpdateParams up, const pfnav_agent *__restrict__ agents,
                const pf_record *__restrict__ rec, const pfnav_movestate *__restrict__ mss,
                const pfnav_flock *__restrict__ flocks, const pf_arrival_dev *__restrict__ arr,
                const float2 *__restrict__ mc_tiles, int nlayers, const uint32_t *__restrict__ work, int nwork,
                const float2 *__restrict__ vel_in, const float2 *__restrict__ vdes_in, pfnav_patch *__restrict__ out,
                const int32_t *__restrict__ flock_of, const pfnav_movestate_ext *__restrict__ exts,
                const pfnav_formation_in *__restrict__ forms, const pf_arrival_dev *__restrict__ tarr,
                const float2 *__restrict__ t_tiles)
{
    const int wi = blockIdx.x * blockDim.x + threadIdx.x;
    if (wi >= nwork) return;
    const uint32_t uid = work[wi];
    const pfnav_agent a = agents[uid];
    const pfnav_movestate ms = mss[uid];
    const float EPS = 1.0f / 1024;
    pfnav_patch p;
    memset(&p, 0, sizeof(p));
    p.next_state = -1;
    p.wait_ticks_left = exts ? exts[uid].wait_ticks_left : 0;
    const quat ms_rot = {ms.next_rot[0], ms.next_rot[1], ms.next_rot[2], ms.next_rot[3]};
    v2 new_vel = {vel_in[wi].x, vel_in[wi].y};
    const v2 vdes = {vdes_in[wi].x, vdes_in[wi].y};
    const v2 ms_vel = {a.velocity[0], a.velocity[1]};
    const v2 curr_xz = {a.pos[0], a.pos[1]};
    const uint32_t state = a.state;

    // flush an unfinished interpolation (movement.c:2311)
    if (ms.left > 0) {
        p.flags |= PFNAV_UPDATE_SET_POSITION | PFNAV_UPDATE_SET_ROTATION | PFNAV_UPDATE_SET_LEFT;
        p.next_pos[0] = ms.next_pos[0]; p.next_pos[1] = ms.next_pos[1]; p.next_pos[2] = ms.next_pos[2];
        p.next_rot[0] = ms_rot.x; p.next_rot[1] = ms_rot.y; p.next_rot[2] = ms_rot.z; p.next_rot[3] = ms_rot.w;
        p.next_left = 0;
    }
    // heading gate (movement.c:2323-2335)
    bool turn_to_move = false;
    quat travel_dir = ms_rot;
    const bool gated = state == PFNAV_STATE_MOVING || state == PFNAV_STATE_SEEK_ENEMIES ||
                       state == PFNAV_STATE_SURROUND_ENTITY || state == PFNAV_STATE_ENTER_ENTITY_RANGE;
    if (v2_len(new_vel) > EPS && gated) {
        travel_dir = dir_quat_from_velocity(v2_len(vdes) > EPS ? vdes : new_vel);
        const float heading_err = (float)fabs((double)quat_pitch_diff(ms_rot, travel_dir) * (180.0f / PF_PI_D));
        const float tolerance = (v2_len(ms_vel) > EPS) ? 90.0f : 10.0f;
        if (heading_err > tolerance) { turn_to_move = true; new_vel = {0.0f, 0.0f}; }
    }
    v2 new_pos_xz = v2_add(curr_xz, new_vel);
    const int layer = nav_layer_for(a.flags, a.radius);
    const bool still = state == PFNAV_STATE_ARRIVED || state == PFNAV_STATE_WAITING;
    if (a.flags & PFNAV_FLAG_GARRISONED) {
        if (!still) { p.flags |= PFNAV_UPDATE_SET_STATE; p.next_state = PFNAV_STATE_ARRIVED; p.next_block = 0; }
        out[wi] = p;
        return;
    }
    bool dummy, on_blocked, np_path, np_blocked;
    probe_tile(m, layer, curr_xz.x, curr_xz.z, dummy, on_blocked);
    probe_tile(m, layer, new_pos_xz.x, new_pos_xz.z, np_path, np_blocked);
    const float turn_rate = up.turn_rate;
    if (v2_len(new_vel) > 0 && np_path && (on_blocked || !np_blocked)) {
        // unit_height (movement.c:2198) with M_HeightAtPoint == 0: terrain height is render state
        const float y = (a.flags & PFNAV_FLAG_WATER) ? 0.0f : (a.flags & PFNAV_FLAG_AIR) ? 20.0f : 0.0f  /* AIR_UNIT_HEIGHT, game.h:50 */;
        p.flags |= PFNAV_UPDATE_SET_PREV_POS | PFNAV_UPDATE_SET_NEXT_POS | PFNAV_UPDATE_SET_STEP | PFNAV_UPDATE_SET_LEFT;
        p.next_ppos[0] = ms.next_pos[0]; p.next_ppos[1] = ms.next_pos[1]; p.next_ppos[2] = ms.next_pos[2];
        p.next_npos[0] = new_pos_xz.x; p.next_npos[1] = y; p.next_npos[2] = new_pos_xz.z;
        p.next_step = 1.0f / (20 / up.hz);
        p.next_left = (float)((20 / up.hz) - 1);
        p.flags |= PFNAV_UPDATE_SET_POSITION;
        if ((20 / up.hz) - 1 == 0) {
            p.next_pos[0] = new_pos_xz.x; p.next_pos[1] = y; p.next_pos[2] = new_pos_xz.z;
        } else {
            // interpolate_positions(next_ppos, next_npos, ms->step) (movement.c:2221)
            float ix, iy, iz;
            if (fabs(1.0 - (double)ms.step) < EPS) { ix = p.next_npos[0]; iy = p.next_npos[1]; iz = p.next_npos[2]; }
            else {
                ix = p.next_ppos[0] + (p.next_npos[0] - p.next_ppos[0]) * ms.step;
                iy = p.next_ppos[1] + (p.next_npos[1] - p.next_ppos[1]) * ms.step;
                iz = p.next_ppos[2] + (p.next_npos[2] - p.next_ppos[2]) * ms.step;
            }
            new_pos_xz = {ix, iz};
            p.next_pos[0] = ix; p.next_pos[1] = iy; p.next_pos[2] = iz;
        }
        p.flags |= PFNAV_UPDATE_SET_VELOCITY;
        p.next_velocity[0] = new_vel.x; p.next_velocity[1] = new_vel.z;
        p.flags |= PFNAV_UPDATE_SET_PREV_ROT | PFNAV_UPDATE_SET_NEXT_ROT | PFNAV_UPDATE_SET_ROTATION;
        p.next_prot[0] = ms_rot.x; p.next_prot[1] = ms_rot.y; p.next_prot[2] = ms_rot.z; p.next_prot[3] = ms_rot.w;
        // orient_to_velocity_history (movement.c:2291) over vel_wma (movement.c:2067)
        v2 wma = {0.0f, 0.0f};
        float denom = 0.0f;
        for (int i = 0; i < PFNAV_VEL_HIST_LEN; i++) {
            const int k = (ms.vel_hist_idx + i) % PFNAV_VEL_HIST_LEN;
            const float wgt = (float)(PFNAV_VEL_HIST_LEN - i);
            wma.x = wma.x + ms.vel_hist[k][0] * wgt;
            wma.z = wma.z + ms.vel_hist[k][1] * wgt;
            denom += wgt;
        }
        if (denom > EPS) { const float inv = 1.0f / denom; wma.x = wma.x * inv; wma.z = wma.z * inv; }
        quat nrot = ms_rot;
        if (v2_len(wma) > EPS) nrot = turn_toward(ms_rot, dir_quat_from_velocity(wma), turn_rate);
        p.next_nrot[0] = nrot.x; p.next_nrot[1] = nrot.y; p.next_nrot[2] = nrot.z; p.next_nrot[3] = nrot.w;
        p.next_rot[0] = ms_rot.x; p.next_rot[1] = ms_rot.y; p.next_rot[2] = ms_rot.z; p.next_rot[3] = ms_rot.w;
    } else {
        p.flags |= PFNAV_UPDATE_SET_VELOCITY;
        p.next_velocity[0] = 0.0f; p.next_velocity[1] = 0.0f;
        const bool held = a.flags & PFNAV_FLAG_COMBAT_HELD;
        if (held || turn_to_move) {
            p.flags |= PFNAV_UPDATE_SET_PREV_ROT | PFNAV_UPDATE_SET_NEXT_ROT | PFNAV_UPDATE_SET_ROTATION |
                       PFNAV_UPDATE_TURNING_IN_PLACE;
            const quat tgt = held ? quat{ms.combat_facing[0], ms.combat_facing[1], ms.combat_facing[2], ms.combat_facing[3]}
                                  : travel_dir;
            const quat nrot = turn_toward(ms_rot, tgt, turn_rate);
            p.next_prot[0] = ms_rot.x; p.next_prot[1] = ms_rot.y; p.next_prot[2] = ms_rot.z; p.next_prot[3] = ms_rot.w;
            p.next_nrot[0] = nrot.x; p.next_nrot[1] = nrot.y; p.next_nrot[2] = nrot.z; p.next_nrot[3] = nrot.w;
            p.next_rot[0] = ms_rot.x; p.next_rot[1] = ms_rot.y; p.next_rot[2] = ms_rot.z; p.next_rot[3] = ms_rot.w;
        }
    }
    // stuck on non-pathable terrain: keep the state (movement.c:2417)
    bool cur_path, cur_blk;
    probe_tile(m, layer, new_pos_xz.x, new_pos_xz.z, cur_path, cur_blk);
    if (!cur_path) { out[wi] = p; return; }
    pfnav_movestate_ext ex;
    if (exts) ex = exts[uid];
    else { memset(&ex, 0, sizeof(ex)); ex.surround_target_uid = PFNAV_NULL_UID; ex.rot[3] = 1.0f; }
    p.wait_ticks_left = ex.wait_ticks_left;
    uint32_t fflags = 0;
    pfnav_formation_in fin;
    if (forms) { fin = forms[uid]; fflags = fin.flags; }
    const bool has_fid = fflags & PFNAV_FORM_HAS_FORMATION;

    switch (state) {
    case PFNAV_STATE_MOVING:
    case PFNAV_STATE_MOVING_IN_FORMATION: {
        // movement.c:2424-2500 (no arrival group: G_ArrivalGroup_ForLayer is the f-1 consumer, inactive here)
        if (has_fid && !(fflags & PFNAV_FORM_ASSIGNMENT_READY)) break;
        if (has_fid && (fflags & PFNAV_FORM_ASSIGNED_TO_CELL) && (fflags & PFNAV_FORM_IN_RANGE_OF_CELL)) {
            p.flags |= PFNAV_UPDATE_SET_STATE; p.next_state = PFNAV_STATE_ARRIVING_TO_CELL;
            break;
        }
        if (a.flock < 0) break;
        const pfnav_flock fl = flocks[a.flock];
        const pf_arrival_dev ac = arr[(size_t)a.flock * nlayers + layer];
        bool arrived = false;
        {   // arrived() (movement.c:2170)
            const v2 tgt = {fl.target[0], fl.target[1]};
            const float thresh = a.radius * 1.5f;
            if (v2_len(v2_sub(tgt, new_pos_xz)) < thresh) arrived = true;
            if (!arrived) {
                // N_IsAdjacentToImpassable (nav.c:4745) && N_IsMaximallyClose (nav.c:4707)
                tile_desc td;
                bool adj = false;
                if (desc_for_point(m, new_pos_xz.x, new_pos_xz.z, td)) {
                    const int ar = td.chunk_r * 64 + td.tile_r, acol = td.chunk_c * 64 + td.tile_c;
                    const int dr[4] = {-1, 0, 0, 1}, dc[4] = {0, -1, 1, 0};
                    for (int e = 0; e < 4 && !adj; e++) {
                        const int nr = ar + dr[e], nc = acol + dc[e];
                        if (nr < 0 || nr >= m.H64 || nc < 0 || nc >= m.W64) continue;
                        adj = tile_blocked_abs(m, layer, nr, nc);
                    }
                }
                if (adj) {
                    for (int i = 0; i < ac.mc_n && !arrived; i++) {
                        const float2 tc = mc_tiles[ac.mc_off + i];
                        const v2 d = {tc.x - new_pos_xz.x, tc.y - new_pos_xz.z};
                        if (v2_len(d) <= thresh) arrived = true;
                    }
                }
            }
            if (!arrived && ac.nearest_ok) {
                const v2 d = {ac.nearest[0] - new_pos_xz.x, ac.nearest[1] - new_pos_xz.z};
                if (v2_len(d) < thresh) arrived = true;
            }
        }
        if (!arrived) {
            // adjacent_flock_members (movement.c:953): any flock member within r + r' + ADJACENCY_SEP_DIST that has
            // ARRIVED. The spatial index pre-selects (radius r + max radius + 5 plus slack); the test itself is
            // the reference's float comparison, so the outcome does not depend on the pre-selection.
            const int32_t icx = bg_scale(curr_xz.x), icy = bg_scale(curr_xz.z), ir = bg_scale(a.radius + up.adj_query_r);
            const int cx_lo = max((icx - ir - g.origin_x) >> 12, 0), cx_hi = min((icx + ir - g.origin_x) >> 12, g.grid_w - 1);
            const int cy_lo = max((icy - ir - g.origin_y) >> 12, 0), cy_hi = min((icy + ir - g.origin_y) >> 12, g.grid_h - 1);
            for (int cy = cy_lo; cy <= cy_hi && !arrived; cy++)
                for (int cx = cx_lo; cx <= cx_hi && !arrived; cx++) {
                    const int c = cy * g.grid_w + cx;
                    const uint32_t b = g.cell_start[c], cnt = g.cell_count[c];
                    for (uint32_t k = 0; k < cnt; k++) {
                        const uint32_t o = g.id[b + k];
                        if (o == uid) continue;
                        const pf_record r = rec[o];
                        if ((r.state_flags >> 24) != PFNAV_STATE_ARRIVED) continue;
                        if (flock_of[o] != a.flock) continue;
                        const v2 d = {curr_xz.x - r.px, curr_xz.z - r.pz};
                        if (v2_len(d) <= a.radius + r.radius + 5.0f) { arrived = true; break; }
                    }
                }
        }
        if (arrived) {
            p.flags |= PFNAV_UPDATE_SET_STATE; p.next_state = PFNAV_STATE_ARRIVED; p.next_block = 1;
        } else if (v2_len(vdes) < EPS) {
            p.flags |= PFNAV_UPDATE_SET_STATE; p.next_state = PFNAV_STATE_WAITING; p.next_block = 1;
        }
        break;
    }
    case PFNAV_STATE_SEEK_ENEMIES:
        break;                              // stays a soft obstacle and retries next tick (movement.c:2501)
    case PFNAV_STATE_SURROUND_ENTITY:
        if (ex.surround_target_uid == PFNAV_NULL_UID) {
            p.flags |= PFNAV_UPDATE_SET_STATE; p.next_state = PFNAV_STATE_ARRIVED; p.next_block = 1;
        } else p.engine_todo |= PFNAV_TODO_SURROUND_QUERY;
        break;
    case PFNAV_STATE_ENTER_ENTITY_RANGE: {
        // movement.c:2569-2603
        if (ex.surround_target_uid == PFNAV_NULL_UID) {
            p.flags |= PFNAV_UPDATE_SET_STATE; p.next_state = PFNAV_STATE_ARRIVED; p.next_block = 1;
            break;
        }
        const pf_record tr = rec[ex.surround_target_uid];
        const v2 xz_target = {tr.px, tr.pz};
        const v2 delta = v2_sub(new_pos_xz, xz_target);
        bool stop = v2_len(delta) <= ex.target_range;
        if (!stop && tarr) {
            // M_NavIsAdjacentToImpassable(new_pos) && M_NavIsMaximallyClose(new_pos, target, 0): the tile list of the
            // target is precomputed per work item on the host (pfnav_arrival_consts with the target's position)
            tile_desc td;
            bool adj = false;
            if (desc_for_point(m, new_pos_xz.x, new_pos_xz.z, td)) {
                const int ar = td.chunk_r * 64 + td.tile_r, acol = td.chunk_c * 64 + td.tile_c;
                const int dr[4] = {-1, 0, 0, 1}, dc[4] = {0, -1, 1, 0};
                for (int e = 0; e < 4 && !adj; e++) {
                    const int nr = ar + dr[e], nc = acol + dc[e];
                    if (nr < 0 || nr >= m.H64 || nc < 0 || nc >= m.W64) continue;
                    adj = tile_blocked_abs(m, layer, nr, nc);
                }
            }
            if (adj) {
                const pf_arrival_dev ta = tarr[wi];
                for (int i = 0; i < ta.mc_n && !stop; i++) {
                    const float2 tc = t_tiles[ta.mc_off + i];
                    const v2 d = {tc.x - new_pos_xz.x, tc.y - new_pos_xz.z};
                    if (v2_len(d) <= 0.0f) stop = true;          // tolerance 0.0f (movement.c:2584)
                }
            }
        }
        if (stop) { p.flags |= PFNAV_UPDATE_SET_STATE; p.next_state = PFNAV_STATE_WAITING; p.next_block = 1; break; }
        const v2 tdelta = v2_sub(xz_target, v2{ex.target_prev_pos[0], ex.target_prev_pos[1]});
        if (v2_len(tdelta) > 5.0f) {
            p.flags |= PFNAV_UPDATE_SET_DEST | PFNAV_UPDATE_SET_TARGET_PREV;
            p.next_dest[0] = xz_target.x; p.next_dest[1] = xz_target.z; p.next_attack = 0;
            p.next_target_prev[0] = xz_target.x; p.next_target_prev[1] = xz_target.z;
        }
        break;
    }
    case PFNAV_STATE_TURNING: {
        // movement.c:2605-2627
        const quat ent_rot = {ex.rot[0], ex.rot[1], ex.rot[2], ex.rot[3]};
        const quat tdir = {ex.target_dir[0], ex.target_dir[1], ex.target_dir[2], ex.target_dir[3]};
        const float degrees = (float)((double)quat_pitch_diff(ent_rot, tdir) * (180.0f / PF_PI_D));
        if (fabs((double)degrees) <= 5.0f) {
            p.flags |= PFNAV_UPDATE_SET_STATE; p.next_state = PFNAV_STATE_ARRIVED; p.next_block = 1;
            break;
        }
        const quat fin_rot = turn_toward(ent_rot, tdir, turn_rate);
        p.flags |= PFNAV_UPDATE_SET_ROTATION | PFNAV_UPDATE_SET_PREV_ROT;
        p.next_rot[0] = fin_rot.x; p.next_rot[1] = fin_rot.y; p.next_rot[2] = fin_rot.z; p.next_rot[3] = fin_rot.w;
        p.next_prot[0] = fin_rot.x; p.next_prot[1] = fin_rot.y; p.next_prot[2] = fin_rot.z; p.next_prot[3] = fin_rot.w;
        break;
    }
    case PFNAV_STATE_WAITING:
        // movement.c:2630-2645: the countdown lives in the movestate itself
        p.wait_ticks_left = ex.wait_ticks_left - 1;
        if (p.wait_ticks_left == 0) { p.flags |= PFNAV_UPDATE_SET_MOVING; p.next_state = ex.wait_prev; }
        break;
    case PFNAV_STATE_ARRIVED:
        break;
    case PFNAV_STATE_ARRIVING_TO_CELL:
        // movement.c:2648-2670
        if (!has_fid) { p.flags |= PFNAV_UPDATE_SET_STATE; p.next_state = PFNAV_STATE_MOVING; break; }
        if (!(fflags & PFNAV_FORM_ASSIGNMENT_READY)) break;
        if (!(fflags & PFNAV_FORM_IN_RANGE_OF_CELL)) { p.flags |= PFNAV_UPDATE_SET_STATE; p.next_state = PFNAV_STATE_MOVING_IN_FORMATION; break; }
        if (fflags & PFNAV_FORM_ARRIVED_AT_CELL) {
            p.flags |= PFNAV_UPDATE_SET_STATE | PFNAV_UPDATE_SET_TARGET_DIR; p.next_state = PFNAV_STATE_TURNING;
            for (int i = 0; i < 4; i++) p.next_target_dir[i] = fin.target_orientation[i];
        }
        break;
    default:
        break;
    }
    {   // ent_update_using_surround_field (movement.c:2672-2691) runs at the end of the apply on the tick's snapshot
        // positions; it is evaluated here, where the snapshot is still intact, and carried in the patch
        const uint32_t after = (p.flags & (PFNAV_UPDATE_SET_STATE | PFNAV_UPDATE_SET_MOVING)) ? (uint32_t)p.next_state : state;
        if (after == PFNAV_STATE_SURROUND_ENTITY && ex.surround_target_uid != PFNAV_NULL_UID) {
            const pf_record tr = rec[ex.surround_target_uid];
            const float dx = (float)fabs((double)(tr.px - curr_xz.x)), dz = (float)fabs((double)(tr.pz - curr_xz.z));
            const float low = (float)(256.0 / 3.0f), high = (float)(256.0 / 2.0f);       // CHUNK_WIDTH / 3, / 2 (movement.c:440-443)
            if (!ex.using_surround_field) { if (dx < low && dz < low) p.engine_todo |= PFNAV_TODO_USE_SURROUND_FIELD; }
            else if (dx >= high || dz >= high) p.engine_todo |= PFNAV_TODO_DROP_SURROUND_FIELD;
        }
    }
    out[wi] = p;
}

// entity_apply_update (movement.c:2693-2757), movestate fields only
__global__ void k_entity_apply(pfnav_agent *__restrict__ agents, pfnav_movestate *__restrict__ mss,
                               pf_record *__restrict__ rec, const uint32_t *__restrict__ work, int nwork,
                               const pfnav_patch *__restrict__ patches, pfnav_movestate_ext *__restrict__ exts)
{
    const int wi = blockIdx.x * blockDim.x + threadIdx.x;
    if (wi >= nwork) return;
    const uint32_t uid = work[wi];
    const pfnav_patch p = patches[wi];
    pfnav_agent a = agents[uid];
    pfnav_movestate ms = mss[uid];
    const float EPS = 1.0f / 1024;
    if (a.flags & PFNAV_FLAG_GARRISONED) return;
    pfnav_movestate_ext ex;
    if (exts) { ex = exts[uid]; ex.wait_ticks_left = p.wait_ticks_left; }      // the countdown of STATE_WAITING (movement.c:2633)
    if (p.flags & PFNAV_UPDATE_SET_STATE) {
        if (p.next_state == PFNAV_STATE_ARRIVED || p.next_state == PFNAV_STATE_WAITING) {
            // entity_finish_moving (movement.c:685): blockers / events / combat stance stay with the engine
            if (exts && p.next_state == PFNAV_STATE_WAITING) { ex.wait_prev = (int32_t)a.state; ex.wait_ticks_left = 60; }     // WAIT_TICKS
            a.velocity[0] = 0.0f; a.velocity[1] = 0.0f;
        }
        a.state = (uint32_t)p.next_state;
    }
    if (p.flags & PFNAV_UPDATE_SET_VELOCITY) {
        a.velocity[0] = p.next_velocity[0]; a.velocity[1] = p.next_velocity[1];
        if (p.flags & PFNAV_UPDATE_TURNING_IN_PLACE) {
            for (int i = 0; i < PFNAV_VEL_HIST_LEN; i++) { ms.vel_hist[i][0] = 0.0f; ms.vel_hist[i][1] = 0.0f; }
        } else {
            bool empty = true;
            for (int i = 0; i < PFNAV_VEL_HIST_LEN; i++)
                if (sqrtf(ms.vel_hist[i][0] * ms.vel_hist[i][0] + ms.vel_hist[i][1] * ms.vel_hist[i][1]) > EPS) empty = false;
            const float vl = sqrtf(a.velocity[0] * a.velocity[0] + a.velocity[1] * a.velocity[1]);
            if (empty && vl > EPS) {
                // seed_vel_hist_facing (movement.c:2046) / facing_dir (:2040)
                const float theta = (float)(2.0 * atan2((double)ms.next_rot[1], (double)ms.next_rot[3]));
                const float dx = (float)(-sin((double)theta)), dz = (float)cos((double)theta);
                for (int i = 0; i < PFNAV_VEL_HIST_LEN; i++) { ms.vel_hist[i][0] = dx * vl; ms.vel_hist[i][1] = dz * vl; }
            }
            ms.vel_hist[ms.vel_hist_idx][0] = a.velocity[0]; ms.vel_hist[ms.vel_hist_idx][1] = a.velocity[1];
            ms.vel_hist_idx = (ms.vel_hist_idx + 1) % PFNAV_VEL_HIST_LEN;
        }
    }
    if (p.flags & PFNAV_UPDATE_SET_POSITION) { a.pos[0] = p.next_pos[0]; a.pos[1] = p.next_pos[2]; }
    if (p.flags & PFNAV_UPDATE_SET_PREV_POS) { a.prev_pos[0] = p.next_ppos[0]; a.prev_pos[1] = p.next_ppos[2]; }
    if (p.flags & PFNAV_UPDATE_SET_NEXT_POS) { ms.next_pos[0] = p.next_npos[0]; ms.next_pos[1] = p.next_npos[1]; ms.next_pos[2] = p.next_npos[2]; }
    if (p.flags & PFNAV_UPDATE_SET_STEP) ms.step = p.next_step;
    if (p.flags & PFNAV_UPDATE_SET_LEFT) ms.left = (int)p.next_left;
    if (p.flags & PFNAV_UPDATE_SET_NEXT_ROT) { ms.next_rot[0] = p.next_nrot[0]; ms.next_rot[1] = p.next_nrot[1]; ms.next_rot[2] = p.next_nrot[2]; ms.next_rot[3] = p.next_nrot[3]; }
    if (exts) {
        if (p.flags & PFNAV_UPDATE_SET_ROTATION) { ex.rot[0] = p.next_rot[0]; ex.rot[1] = p.next_rot[1]; ex.rot[2] = p.next_rot[2]; ex.rot[3] = p.next_rot[3]; }   // Entity_SetRot
        if (p.flags & PFNAV_UPDATE_SET_TARGET_PREV) { ex.target_prev_pos[0] = p.next_target_prev[0]; ex.target_prev_pos[1] = p.next_target_prev[1]; }
        if (p.flags & PFNAV_UPDATE_SET_TARGET_DIR) { ex.target_dir[0] = p.next_target_dir[0]; ex.target_dir[1] = p.next_target_dir[1]; ex.target_dir[2] = p.next_target_dir[2]; ex.target_dir[3] = p.next_target_dir[3]; }
    }
    if (p.flags & PFNAV_UPDATE_SET_MOVING) {
        // movement.c:2750-2755: entity_unblock is the engine's; move_notify_motion_start wipes the velocity history
        if (!(a.flags & PFNAV_FLAG_COMBAT_HELD))
            for (int i = 0; i < PFNAV_VEL_HIST_LEN; i++) { ms.vel_hist[i][0] = 0.0f; ms.vel_hist[i][1] = 0.0f; }
        a.state = (uint32_t)p.next_state;
    }
    if (exts && a.state == PFNAV_STATE_SURROUND_ENTITY) {
        // ent_update_using_surround_field (movement.c:2672), decided by the update pass on the tick's snapshot positions
        if (p.engine_todo & PFNAV_TODO_USE_SURROUND_FIELD) ex.using_surround_field = 1;
        if (p.engine_todo & PFNAV_TODO_DROP_SURROUND_FIELD) ex.using_surround_field = 0;
    }
    if (exts) exts[uid] = ex;
    agents[uid] = a;
    mss[uid] = ms;
    pf_record r;
    r.px = a.pos[0]; r.pz = a.pos[1]; r.vx = a.velocity[0]; r.vz = a.velocity[1];
    r.radius = a.radius;
    r.state_flags = (a.state << 24) | (a.flags & 0xFFFFFFu);
    rec[uid] = r;
}

extern "C" int pfnav_agents_upload_movestate(pfnav_ctx *ctx, const pfnav_movestate *ms, size_t n)
{
    PF_ARG(ctx && ctx->d_agents && ms, "agents not uploaded / ms == NULL");
    PF_ARG(n == ctx->shard_hi - ctx->shard_lo, "one movestate record per uploaded agent (of this context's own range)");
    PF_CUDA(cudaSetDevice(ctx->device));
    int rc;
    if (ctx->n_agents > ctx->cap_movestate) {
        if ((rc = ensure(ctx->d_movestate, ctx->cap_movestate, ctx->n_agents))) return rc;
        ctx->cap_movestate = ctx->n_agents;
    }
    PF_CUDA(cudaMemcpy(ctx->d_movestate + ctx->shard_lo, ms, n * sizeof(pfnav_movestate), cudaMemcpyHostToDevice));
    ctx->movestate_set = true;
    return PFNAV_OK;
}

extern "C" int pfnav_agents_upload_formation(pfnav_ctx *ctx, const pfnav_formation_in *f, size_t n)
{
    PF_ARG(ctx && ctx->d_agents && f, "agents not uploaded / null");
    PF_ARG(n == ctx->shard_hi - ctx->shard_lo, "one record per uploaded agent (of this context's own range)");
    PF_CUDA(cudaSetDevice(ctx->device));
    if (ctx->n_agents > ctx->cap_formation) {
        cudaFree(ctx->d_formation); ctx->d_formation = nullptr; ctx->cap_formation = 0;
        PF_CUDA(cudaMalloc(&ctx->d_formation, ctx->n_agents * sizeof(pfnav_formation_in)));
        PF_CUDA(cudaMemset(ctx->d_formation, 0, ctx->n_agents * sizeof(pfnav_formation_in)));
        ctx->cap_formation = ctx->n_agents;
    }
    PF_CUDA(cudaMemcpy(ctx->d_formation + ctx->shard_lo, f, n * sizeof(pfnav_formation_in), cudaMemcpyHostToDevice));
    return PFNAV_OK;
}

extern "C" int pfnav_agents_upload_movestate_ext(pfnav_ctx *ctx, const pfnav_movestate_ext *ms, size_t n)
{
    PF_ARG(ctx && ctx->d_agents && ms, "agents not uploaded / null");
    PF_ARG(n == ctx->shard_hi - ctx->shard_lo, "one record per uploaded agent (of this context's own range)");
    PF_CUDA(cudaSetDevice(ctx->device));
    if (ctx->n_agents > ctx->cap_ms_ext) {
        cudaFree(ctx->d_ms_ext); ctx->d_ms_ext = nullptr; ctx->cap_ms_ext = 0;
        PF_CUDA(cudaMalloc(&ctx->d_ms_ext, ctx->n_agents * sizeof(pfnav_movestate_ext)));
        ctx->cap_ms_ext = ctx->n_agents;
    }
    PF_CUDA(cudaMemcpy(ctx->d_ms_ext + ctx->shard_lo, ms, n * sizeof(pfnav_movestate_ext), cudaMemcpyHostToDevice));
    return PFNAV_OK;
}

// work items in STATE_ENTER_ENTITY_RANGE and where their target stands (for the N_IsMaximallyClose clause, movement.c:2583)
struct pf_enter_item { uint32_t wi; int32_t layer; float tx, tz; };
__global__ void k_collect_enter_range(const pfnav_agent *__restrict__ agents, const pfnav_movestate_ext *__restrict__ exts,
                                      const pf_record *__restrict__ rec, const uint32_t *__restrict__ work, int nwork,
                                      pf_enter_item *__restrict__ out, uint32_t *__restrict__ count)
{
    const int wi = blockIdx.x * blockDim.x + threadIdx.x;
    if (wi >= nwork) return;
    const uint32_t uid = work[wi];
    const pfnav_agent a = agents[uid];
    if (a.state != PFNAV_STATE_ENTER_ENTITY_RANGE || exts[uid].surround_target_uid == PFNAV_NULL_UID) return;
    const pf_record tr = rec[exts[uid].surround_target_uid];
    const uint32_t k = atomicAdd(count, 1u);
    out[k] = {(uint32_t)wi, nav_layer_for(a.flags, a.radius), tr.px, tr.pz};
}

static int refresh_arrival_consts(pfnav_ctx *ctx, cudaStream_t st)
{
    if (ctx->arrival_valid && ctx->arrival_epoch == ctx->map_epoch) return 0;
    const size_t nf = ctx->n_flocks, nl = (size_t)ctx->nlayers;
    std::vector<pf_arrival_dev> dev(nf * nl);
    std::vector<float2> tiles;
    pf_arrival_consts c;
    for (size_t f = 0; f < nf; f++)
        for (size_t l = 0; l < nl; l++) {
            pf_arrival_dev &d = dev[f * nl + l];
            memset(&d, 0, sizeof(d));
            if (!ctx->flock_layer_used[f * PFNAV_NAV_LAYER_MAX + l]) continue;
            int rc = pfnav_arrival_consts(ctx, (int)l, ctx->h_flocks[f].target[0], ctx->h_flocks[f].target[1], &c);
            if (rc) return rc;
            d.nearest_ok = c.nearest_ok; d.nearest[0] = c.nearest[0]; d.nearest[1] = c.nearest[1];
            d.mc_n = c.mc_n; d.mc_off = (int32_t)tiles.size();
            for (int i = 0; i < c.mc_n; i++) tiles.push_back(make_float2(c.mc[i][0], c.mc[i][1]));
        }
    const size_t b0 = dev.size() * sizeof(pf_arrival_dev), b1 = std::max<size_t>(tiles.size(), 1) * sizeof(float2);
    if (b0 + b1 > ctx->arrival_bytes) {
        cudaFree(ctx->d_arrival); ctx->d_arrival = nullptr; ctx->arrival_bytes = 0;
        PF_CUDA(cudaMalloc(&ctx->d_arrival, b0 + b1));
        ctx->arrival_bytes = b0 + b1;
    }
    PF_CUDA(cudaStreamSynchronize(st));
    if (b0) PF_CUDA(cudaMemcpy(ctx->d_arrival, dev.data(), b0, cudaMemcpyHostToDevice));
    if (!tiles.empty()) PF_CUDA(cudaMemcpy((uint8_t *)ctx->d_arrival + b0, tiles.data(), tiles.size() * sizeof(float2), cudaMemcpyHostToDevice));
    ctx->arrival_valid = true; ctx->arrival_epoch = ctx->map_epoch;
    return 0;
}

extern "C" int pfnav_agents_compute_updates(pfnav_ctx *ctx, void *stream)
{
    PF_ARG(ctx && ctx->d_agents, "agents not uploaded");
    PF_ARG(ctx->movestate_set, "pfnav_agents_upload_movestate not called");
    if (ctx->n_work == 0) return PFNAV_OK;
    PF_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = pf_stream(ctx, stream);
    int rc;
    if (ctx->n_work > ctx->cap_patches) {
        if ((rc = ensure(ctx->d_patches, ctx->cap_patches, ctx->n_work))) return rc;
        ctx->cap_patches = ctx->n_work;
    }
    if ((rc = refresh_arrival_consts(ctx, st))) return rc;
    if (!ctx->update_done) PF_CUDA(cudaEventCreateWithFlags(&ctx->update_done, cudaEventDisableTiming));
    MapView m;
    m.cost = ctx->d_cost; m.blk = ctx->d_blk; m.W64 = ctx->W64; m.H64 = ctx->H64;
    m.chunk_w = ctx->chunk_w; m.chunk_h = ctx->chunk_h; m.map_x = ctx->map_x; m.map_z = ctx->map_z;
    UpdateParams up;
    up.hz = ctx->hz;
    up.turn_rate = (float)((double)(15.0f / (float)ctx->hz) * 20.0);      // SCALED_MAX_TURN_RATE (movement.c:434)
    up.adj_query_r = ctx->max_radius + 5.0f + 0.0625f;
    const int nwork = (int)ctx->n_work;
    const size_t b0 = ctx->n_flocks * (size_t)ctx->nlayers * sizeof(pf_arrival_dev);
    // STATE_ENTER_ENTITY_RANGE: the tiles N_IsMaximallyClose (nav.c:4707) compares with depend on where the TARGET entity
    // stands; they are computed on the host per work item in that state (a handful of units at a time)
    const pf_arrival_dev *d_tarr = nullptr; const float2 *d_ttiles = nullptr;
    if (ctx->d_ms_ext) {
        if ((size_t)nwork > ctx->cap_enter) {
            cudaFree(ctx->d_enter); ctx->d_enter = nullptr; ctx->cap_enter = 0;
            PF_CUDA(cudaMalloc(&ctx->d_enter, (size_t)nwork * (sizeof(pf_enter_item) + sizeof(pf_arrival_dev)) + 16));
            ctx->cap_enter = (size_t)nwork;
        }
        uint32_t *d_cnt = (uint32_t *)ctx->d_enter;
        pf_enter_item *d_items = (pf_enter_item *)((uint8_t *)ctx->d_enter + 16);
        pf_arrival_dev *d_t = (pf_arrival_dev *)(d_items + nwork);
        PF_CUDA(cudaMemsetAsync(d_cnt, 0, 4, st));
        k_collect_enter_range<<<(nwork + 127) / 128, 128, 0, st>>>(ctx->d_agents, ctx->d_ms_ext, ctx->d_records, ctx->d_work, nwork, d_items, d_cnt);
        ctx->launches++;
        uint32_t cnt = 0;
        PF_CUDA(cudaMemcpyAsync(&cnt, d_cnt, 4, cudaMemcpyDeviceToHost, st));
        PF_CUDA(cudaStreamSynchronize(st));
        if (cnt) {
            std::vector<pf_enter_item> items(cnt);
            PF_CUDA(cudaMemcpy(items.data(), d_items, cnt * sizeof(pf_enter_item), cudaMemcpyDeviceToHost));
            std::vector<pf_arrival_dev> tarr(nwork);
            memset(tarr.data(), 0, tarr.size() * sizeof(pf_arrival_dev));
            std::vector<float2> tiles;
            pf_arrival_consts c;
            for (const pf_enter_item &it : items) {
                if ((rc = pfnav_arrival_consts(ctx, it.layer, it.tx, it.tz, &c))) return rc;
                tarr[it.wi].mc_n = c.mc_n; tarr[it.wi].mc_off = (int32_t)tiles.size();
                for (int i = 0; i < c.mc_n; i++) tiles.push_back(make_float2(c.mc[i][0], c.mc[i][1]));
            }
            if (tiles.size() * sizeof(float2) > ctx->cap_ttiles) {
                cudaFree(ctx->d_ttiles); ctx->d_ttiles = nullptr; ctx->cap_ttiles = 0;
                PF_CUDA(cudaMalloc(&ctx->d_ttiles, tiles.size() * sizeof(float2) * 2));
                ctx->cap_ttiles = tiles.size() * sizeof(float2) * 2;
            }
            PF_CUDA(cudaMemcpy(d_t, tarr.data(), tarr.size() * sizeof(pf_arrival_dev), cudaMemcpyHostToDevice));
            if (!tiles.empty()) PF_CUDA(cudaMemcpy(ctx->d_ttiles, tiles.data(), tiles.size() * sizeof(float2), cudaMemcpyHostToDevice));
            d_tarr = d_t; d_ttiles = (const float2 *)ctx->d_ttiles;
        }
    }
    pf_prof_scope prof(ctx, st, PF_PROF_UPDATE);
    k_entity_update<<<(nwork + 127) / 128, 128, 0, st>>>(m, grid_of(ctx), up, ctx->d_agents, ctx->d_records, ctx->d_movestate,
        ctx->d_flocks, (const pf_arrival_dev *)ctx->d_arrival, (const float2 *)((const uint8_t *)ctx->d_arrival + b0),
        ctx->nlayers, ctx->d_work, nwork, ctx->d_vel_out, ctx->d_vdes_out, ctx->d_patches, ctx->d_flock_of,
        ctx->d_ms_ext, ctx->d_formation, d_tarr, d_ttiles);
    ctx->launches++;
    PF_CUDA(cudaGetLastError());
    PF_CUDA(cudaEventRecord(ctx->update_done, st));
    return PFNAV_OK;
}

extern "C" int pfnav_agents_read_patches(pfnav_ctx *ctx, pfnav_patch *out, size_t maxout)
{
    PF_ARG(ctx && out && ctx->d_patches && ctx->update_done, "no update pass has run");
    PF_CUDA(cudaSetDevice(ctx->device));
    PF_CUDA(cudaEventSynchronize(ctx->update_done));
    const size_t n = std::min(maxout, ctx->n_work);
    PF_CUDA(cudaMemcpy(out, ctx->d_patches, n * sizeof(pfnav_patch), cudaMemcpyDeviceToHost));
    return PFNAV_OK;
}

extern "C" int pfnav_agents_apply_updates(pfnav_ctx *ctx, void *stream)
{
    PF_ARG(ctx && ctx->d_patches && ctx->update_done, "no update pass has run");
    if (ctx->n_work == 0) return PFNAV_OK;
    PF_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = pf_stream(ctx, stream);
    const int nwork = (int)ctx->n_work;
    pf_prof_scope prof(ctx, st, PF_PROF_APPLY);
    k_entity_apply<<<(nwork + 127) / 128, 128, 0, st>>>(ctx->d_agents, ctx->d_movestate, ctx->d_records, ctx->d_work, nwork,
                                                       ctx->d_patches, ctx->d_ms_ext);
    ctx->launches++;
    PF_CUDA(cudaGetLastError());
    PF_CUDA(cudaEventRecord(ctx->update_done, st));
    return PFNAV_OK;
}

extern "C" int pfnav_agents_read_state(pfnav_ctx *ctx, pfnav_agent *agents_out, pfnav_movestate *ms_out, size_t maxout)
{
    PF_ARG(ctx && ctx->d_agents, "agents not uploaded");
    PF_CUDA(cudaSetDevice(ctx->device));
    PF_CUDA(cudaDeviceSynchronize());
    const size_t n = std::min(maxout, ctx->shard_hi - ctx->shard_lo);       // this context's own entity range
    if (agents_out && n) PF_CUDA(cudaMemcpy(agents_out, ctx->d_agents + ctx->shard_lo, n * sizeof(pfnav_agent), cudaMemcpyDeviceToHost));
    if (ms_out) {
        PF_ARG(ctx->movestate_set, "no movestate uploaded");
        if (n) PF_CUDA(cudaMemcpy(ms_out, ctx->d_movestate + ctx->shard_lo, n * sizeof(pfnav_movestate), cudaMemcpyDeviceToHost));
    }
    return PFNAV_OK;
}

// ------------------------------------------------------------------------------------------
// On-miss chain of N_DesiredPointSeekVelocity (nav.c:3484-3554) against the device field pool.
// ------------------------------------------------------------------------------------------
struct pf_miss { uint32_t wi, uid; int32_t dest, chunk, tile, liid; float px, pz; };

// work agents that would take one of the on-miss branches of the navigation tick:
//   kind 0  compute_los_state -> N_HasDestLOS (nav.c:4026): no LOS field for (dest, chunk of prev_pos)
//   kind 1  compute_desired_velocity -> N_DesiredPointSeekVelocity (nav.c:3468): no flow field for (dest, chunk of
//           pos), or the agent's own tile has no direction (dir_idx == FD_NONE)
__global__ void k_collect_misses(MapView m, PoolView pool, const uint16_t *__restrict__ liid_img,
                                 const pfnav_agent *__restrict__ agents, const pfnav_flock *__restrict__ flocks,
                                 const uint32_t *__restrict__ work, int nwork, int kind, pf_miss *__restrict__ out,
                                 uint32_t *__restrict__ count, uint32_t cap, const pfnav_movestate_ext *__restrict__ exts, int nlayers_)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= nwork) return;
    const uint32_t uid = work[w];
    const pfnav_agent a = agents[uid];
    if (kind == 2) {
        // kind 2: N_DesiredEnemySeekVelocity / N_DesiredSurroundVelocity (nav.c:3603, 3687): the entity's TARGET_ENEMIES /
        // TARGET_ENTITY field exists but its own tile has no direction
        const bool seek = a.state == PFNAV_STATE_SEEK_ENEMIES;
        const bool surr = a.state == PFNAV_STATE_SURROUND_ENTITY && exts && exts[uid].surround_target_uid != PFNAV_NULL_UID &&
                          exts[uid].using_surround_field;
        const int dest = (int)a.aux_dest1 - 1;
        if (!(seek || surr) || dest < 0 || dest >= pool.ndests) return;
        tile_desc t;
        if (!desc_for_point(m, a.pos[0], a.pos[1], t)) return;
        const int chunks = m.chunk_w * m.chunk_h, chunk = t.chunk_r * m.chunk_w + t.chunk_c;
        const int sl = pool.slot[(size_t)dest * chunks + chunk];
        if (sl < 0 || !(pool.has[sl] & 1)) return;                      // building it needs the engine's entity list
        const int layer = nav_layer_for(a.flags, a.radius);
        const uint16_t li = layer < nlayers_ ? liid_img[((size_t)layer * m.H64 + t.chunk_r * 64 + t.tile_r) * m.W64 + t.chunk_c * 64 + t.tile_c] : 0xFFFF;
        // the surround variant repairs a blocked tile whether or not it already has a direction (nav.c:3729)
        if ((pool.flow[(size_t)sl * 4096 + t.tile_r * 64 + t.tile_c] & 0xF) != 0 && !(surr && li == 0xFFFF)) return;
        const uint32_t k = atomicAdd(count, 1u);
        if (k >= cap) return;
        pf_miss r;
        r.wi = (uint32_t)w; r.uid = uid; r.dest = dest; r.chunk = chunk; r.tile = t.tile_r * 64 + t.tile_c; r.liid = li;
        r.px = a.pos[0]; r.pz = a.pos[1];
        out[k] = r;
        return;
    }
    if (a.flock < 0) return;
    const pfnav_flock fl = flocks[a.flock];
    if (fl.dest < 0 || fl.dest >= pool.ndests) return;
    const float px = kind == 0 ? a.prev_pos[0] : a.pos[0], pz = kind == 0 ? a.prev_pos[1] : a.pos[1];
    tile_desc t;
    if (!desc_for_point(m, px, pz, t)) return;
    const int chunks = m.chunk_w * m.chunk_h, chunk = t.chunk_r * m.chunk_w + t.chunk_c;
    const int s = pool.slot[(size_t)fl.dest * chunks + chunk];
    bool miss;
    if (kind == 0) miss = (s < 0) || !(pool.has[s] & 2);
    else {
        miss = (s < 0) || !(pool.has[s] & 1);
        if (!miss) miss = (pool.flow[(size_t)s * 4096 + t.tile_r * 64 + t.tile_c] & 0xF) == 0;
    }
    if (!miss) return;
    const uint32_t k = atomicAdd(count, 1u);
    if (k >= cap) return;
    pf_miss r;
    r.wi = (uint32_t)w; r.uid = uid; r.dest = fl.dest; r.chunk = chunk; r.tile = t.tile_r * 64 + t.tile_c;
    r.liid = liid_img[((size_t)fl.layer * m.H64 + t.chunk_r * 64 + t.tile_r) * m.W64 + t.chunk_c * 64 + t.tile_c];
    r.px = px; r.pz = pz;
    out[k] = r;
}

// The on-miss branches of one navigation tick against the device pool, in the reference's order: first every
// work item's LOS state (compute_los_state, movement.c:4129: a (dest, chunk) without LOS field requests the path
// from the entity's previous position), then every work item's desired velocity (compute_desired_velocity, :4163:
// request from the entity's position, then the in-place field repair if its tile still has no direction). The
// reference walks the entities serially and every request changes what the next entity finds; here the candidates
// are collected on the device, reduced to one representative per (dest, chunk[, local island | blocked tile]) in work
// order, and each representative is RE-CHECKED against the pool as the requests before it left it -- an entity the
// reference would have found served by an earlier entity's request does not request again.
extern "C" int pfnav_pool_repair(pfnav_ctx *ctx, int *out_nrequests, int *out_nrepairs)
{
    PF_ARG(ctx && ctx->d_agents && ctx->d_pool_slot, "agents / pool missing");
    if (out_nrequests) *out_nrequests = 0;
    if (out_nrepairs) *out_nrepairs = 0;
    if (ctx->n_work == 0) return PFNAV_OK;
    PF_CUDA(cudaSetDevice(ctx->device));
    PF_CUDA(pf_fields_sync(ctx));
    cudaStream_t st = ctx->tick_stream;
    PF_CUDA(cudaDeviceSynchronize());
    const int nwork = (int)ctx->n_work;
    const int chunks = ctx->chunk_w * ctx->chunk_h;
    pf_miss *d_miss = nullptr; uint32_t *d_cnt = nullptr;
    PF_CUDA(cudaMalloc(&d_miss, (size_t)nwork * sizeof(pf_miss)));
    if (cudaMalloc(&d_cnt, 4) != cudaSuccess) { cudaFree(d_miss); pfnav_set_error("cudaMalloc"); return PFNAV_ERR_NOMEM; }
    MapView m;
    m.cost = ctx->d_cost; m.blk = ctx->d_blk; m.W64 = ctx->W64; m.H64 = ctx->H64;
    m.chunk_w = ctx->chunk_w; m.chunk_h = ctx->chunk_h; m.map_x = ctx->map_x; m.map_z = ctx->map_z;
    std::vector<int> flock_of_dest(ctx->pool_ndests, -1);
    for (size_t f = 0; f < ctx->h_flocks.size(); f++)
        if (ctx->h_flocks[f].dest >= 0 && ctx->h_flocks[f].dest < ctx->pool_ndests) flock_of_dest[ctx->h_flocks[f].dest] = (int)f;
    int nreq = 0, nrep = 0, rc = 0;
    auto collect = [&](int kind, std::vector<pf_miss> &reps) -> int {
        PoolView pv;
        pv.slot = ctx->d_pool_slot; pv.flow = ctx->d_pool_flow; pv.los = ctx->d_pool_los;
        pv.has = ctx->d_pool_los + (size_t)ctx->pool_max * 4096; pv.ndests = ctx->pool_ndests;
        pv.touch = nullptr; pv.tick_no = 0;
        PF_CUDA(cudaMemsetAsync(d_cnt, 0, 4, st));
        k_collect_misses<<<(nwork + 127) / 128, 128, 0, st>>>(m, pv, ctx->d_liid, ctx->d_agents, ctx->d_flocks, ctx->d_work, nwork,
                                                             kind, d_miss, d_cnt, (uint32_t)nwork, ctx->d_ms_ext, ctx->nlayers);
        ctx->launches++;
        uint32_t cnt = 0;
        PF_CUDA(cudaMemcpyAsync(&cnt, d_cnt, 4, cudaMemcpyDeviceToHost, st));
        PF_CUDA(cudaStreamSynchronize(st));
        std::vector<pf_miss> miss(std::min<uint32_t>(cnt, (uint32_t)nwork));
        if (!miss.empty()) PF_CUDA(cudaMemcpy(miss.data(), d_miss, miss.size() * sizeof(pf_miss), cudaMemcpyDeviceToHost));
        std::sort(miss.begin(), miss.end(), [](const pf_miss &a, const pf_miss &b) { return a.wi < b.wi; });
        std::vector<uint64_t> seen;
        reps.clear();
        for (const pf_miss &r : miss) {
            uint64_t key = ((uint64_t)r.dest << 40) | ((uint64_t)r.chunk << 20);
            if (kind >= 1) key |= (r.liid == 0xFFFF ? (0x10000u | (uint32_t)r.tile) : (uint32_t)r.liid);
            if (std::find(seen.begin(), seen.end(), key) != seen.end()) continue;
            seen.push_back(key);
            reps.push_back(r);
        }
        return 0;
    };
    auto request_from = [&](const pf_miss &r, int *ok) -> int {
        const int f = flock_of_dest[r.dest];
        *ok = 0;
        if (f < 0) return 0;
        const pfnav_flock &fl = ctx->h_flocks[f];
        int nf = 0, nl = 0; uint32_t did = 0;
        int rc2 = pfnav_pool_request_path(ctx, r.dest, fl.layer, r.px, r.pz, fl.target[0], fl.target[1], nullptr, &did, ok, &nf, &nl);
        if (rc2) return rc2;
        if (nf + nl > 0) nreq++;
        return 0;
    };
    std::vector<pf_miss> reps;
    // ---- LOS state (N_HasDestLOS, nav.c:4038-4045) ----
    if (!(rc = collect(0, reps))) {
        for (const pf_miss &r : reps) {
            const int slot = ctx->h_pool_slot[(size_t)r.dest * chunks + r.chunk];
            if (slot >= 0 && (ctx->h_pool_has[slot] & 2)) continue;          // an earlier entity's request built it
            int ok;
            if ((rc = request_from(r, &ok))) break;
        }
    }
    // ---- desired velocity (N_DesiredPointSeekVelocity, nav.c:3483-3554) ----
    if (!rc && !(rc = collect(1, reps))) {
        for (const pf_miss &r : reps) {
            auto dir_now = [&](int *dir) -> int {          // the tile's direction as the pool holds it right now, -1 = no field
                const int slot = ctx->h_pool_slot[(size_t)r.dest * chunks + r.chunk];
                *dir = -1;
                if (slot < 0 || !(ctx->h_pool_has[slot] & 1)) return 0;
                uint8_t d = 0;
                PF_CUDA(cudaDeviceSynchronize());
                PF_CUDA(cudaMemcpy(&d, ctx->d_pool_flow + (size_t)slot * 4096 + r.tile, 1, cudaMemcpyDeviceToHost));
                *dir = d & 0xF;
                return 0;
            };
            int dir;
            if ((rc = dir_now(&dir))) break;
            if (dir > 0) continue;                                        // served by an earlier entity's request / repair
            int ok;
            if ((rc = request_from(r, &ok))) break;
            if (!ok) continue;                                            // no path from here: zero vector, nothing repaired (nav.c:3501)
            if ((rc = dir_now(&dir))) break;
            if (dir != 0) continue;                                       // case 1: the path query fixed it (or no field: no path)
            const int slot = ctx->h_pool_slot[(size_t)r.dest * chunks + r.chunk];
            const pfnav_field_req tg = ctx->h_pool_req[slot];
            int32_t kind, arg;
            if (r.liid == 0xFFFF) { kind = PFNAV_REPAIR_NEAREST_PATHABLE; arg = ((r.tile >> 6) << 8) | (r.tile & 63); }
            else                  { kind = PFNAV_REPAIR_ISLAND_TO_NEAREST; arg = r.liid; }
            const int32_t sl = slot;
            if ((rc = pfnav_flow_repair_pool(ctx, &tg, &kind, &arg, &sl, 1))) break;
            nrep++;
        }
    }
    // ---- enemy-seek / surround fields: in-place repairs only (nav.c:3647-3675, 3729-3757) ----
    if (!rc && !(rc = collect(2, reps))) {
        for (const pf_miss &r : reps) {
            const int slot = ctx->h_pool_slot[(size_t)r.dest * chunks + r.chunk];
            if (slot < 0 || !(ctx->h_pool_has[slot] & 1)) continue;
            uint8_t d = 0;
            if (cudaDeviceSynchronize() != cudaSuccess ||
                cudaMemcpy(&d, ctx->d_pool_flow + (size_t)slot * 4096 + r.tile, 1, cudaMemcpyDeviceToHost) != cudaSuccess) {
                pfnav_set_error("pfnav_pool_repair: read-back failed"); rc = PFNAV_ERR_CUDA; break;
            }
            const pfnav_field_req tg = ctx->h_pool_req[slot];
            if ((tg.target_type & 0xFF) < 2) continue;
            const bool surround = (tg.target_type & 0xFF) == 2 + PFNAV_TARGET_ENTITY;
            int32_t kind, arg;
            if (r.liid == 0xFFFF) {
                if ((d & 0xF) != 0 && !surround) continue;
                kind = PFNAV_REPAIR_NEAREST_PATHABLE; arg = ((r.tile >> 6) << 8) | (r.tile & 63);
            } else {
                if ((d & 0xF) != 0) continue;
                kind = PFNAV_REPAIR_ISLAND_TO_NEAREST; arg = r.liid;
            }
            const int32_t sl = slot;
            if ((rc = pfnav_flow_repair_pool(ctx, &tg, &kind, &arg, &sl, 1))) break;
            nrep++;
        }
    }
    cudaFree(d_miss); cudaFree(d_cnt);
    if (rc) return rc;
    if (nrep) ctx->goal_batch.valid = false;      // pool bytes changed under a resident plan
    if (out_nrequests) *out_nrequests = nreq;
    if (out_nrepairs) *out_nrepairs = nrep;
    return PFNAV_OK;
}

// Test / tuning hook: 0 = always the single-pass velocity kernel, 1 (default) = split it around the field join
// whenever LOS chains are still in flight, 2 = always split (exercises the two-phase path without fields).
// Test / tuning hook: cohesion pass, 0 = automatic (windowed for flocks of >= 20 000), 1 = always windowed, 2 = always the
// full member list. Results agree to float rounding (the windowed pass falls back per entity where they would not).
extern "C" int pfnav_set_cohesion_mode(pfnav_ctx *ctx, int mode)
{
    PF_ARG(ctx && mode >= 0 && mode <= 2, "mode");
    ctx->cohesion_mode = mode;
    return PFNAV_OK;
}

// Test / tuning hook: how often G_ClearPath_NewVelocity's retry loop (clearpath.c:702-713) was entered since the last reset.
// out[0] = first solves without an admissible velocity, out[1] = of those with both neighbour lists non-empty (the loop
// can run), out[2] = entities that replayed the loop literally (order-dependent tie), out[3] = solves inside those replays.
extern "C" int pfnav_agents_clearpath_stats(pfnav_ctx *ctx, uint64_t *out4, int reset)
{
    PF_ARG(ctx && out4, "args");
    PF_NEED_DEVICE(ctx);
    memset(out4, 0, 4 * sizeof(uint64_t));
    if (!ctx->d_cp_stats) return PFNAV_OK;
    PF_CUDA(cudaSetDevice(ctx->device));
    PF_CUDA(cudaDeviceSynchronize());
    PF_CUDA(cudaMemcpy(out4, ctx->d_cp_stats, 4 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    if (reset) PF_CUDA(cudaMemset(ctx->d_cp_stats, 0, 8 * sizeof(unsigned long long)));
    return PFNAV_OK;
}

extern "C" int pfnav_set_two_phase(pfnav_ctx *ctx, int mode)
{
    PF_ARG(ctx && mode >= 0 && (mode & 3) <= 2, "mode");
    ctx->two_phase = (mode & 3) != 0;
    ctx->two_phase_force = (mode & 3) == 2;
    ctx->phase_a_ctas_per_sm = mode >> 4;          // tuning: bits 4.. = resident CTAs per SM of phase A (0 = default)
    return PFNAV_OK;
}
