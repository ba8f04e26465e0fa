// pfnav_internal.cuh -- shared declarations of libpfnav.so (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <set>
#include <string>
#include <utility>
#include <vector>
#include "../../include/pfnav.h"

#define PFNAV_VERSION 100

void pfnav_set_error(const char *fmt, ...);

#define PF_CUDA(call)                                                                       \
    do {                                                                                    \
        cudaError_t _e = (call);                                                            \
        if (_e != cudaSuccess) {                                                            \
            pfnav_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call,                   \
                            cudaGetErrorString(_e));                                        \
            return PFNAV_ERR_CUDA;                                                          \
        }                                                                                   \
    } while (0)

#define PF_NEED_DEVICE(ctx)                                                                         \
    do {                                                                                            \
        if ((ctx)->device < 0) {                                                                    \
            pfnav_set_error("%s: host-only context has no compute path (no CPU fallback)", __func__); \
            return PFNAV_ERR_NO_DEVICE;                                                             \
        }                                                                                           \
    } while (0)

#define PF_ARG(cond, msg)                                                                   \
    do {                                                                                    \
        if (!(cond)) {                                                                      \
            pfnav_set_error("%s:%d: bad argument: %s", __FILE__, __LINE__, msg);            \
            return PFNAV_ERR_ARG;                                                           \
        }                                                                                   \
    } while (0)

// 24-byte neighbour record: exactly what one agent needs to know about another
// (SURVEY.md 8e). This is also the unit of the per-tick NCCL all-gather.
struct __align__(8) pf_record {
    float    px, pz;
    float    vx, vz;
    float    radius;
    uint32_t state_flags;     // state << 24 | (flags & 0xFFFFFF)
};
static_assert(sizeof(pf_record) == 24, "pf_record must be 24 bytes");

struct pfnav_ctx {
    int device = 0;
    int sm_count = 0;
    uint64_t launches = 0;
    bool use_tma = true;
    bool tma_ok = false;

    // ---- map state (row-major images per layer: [layer][H*64][W*64]) ----
    int chunk_w = 0, chunk_h = 0, nlayers = 0;
    int W64 = 0, H64 = 0;
    float map_x = 0.f, map_z = 0.f;
    uint8_t  *d_cost = nullptr;      // u8
    uint16_t *d_blk = nullptr;       // u16 blockers refcounts
    uint16_t *d_liid = nullptr;      // u16 local islands
    // per-faction blocker refcounts (chunk->factions, nav_data.h): host counts per layer (allocated on first
    // use, [chunk][15][4096]) and a device bit mask per tile (bit f <=> factions[f] > 0) for the "attacking"
    // passability rule field_tile_passable_no_enemies (field.c:179)
    uint16_t *d_fmask = nullptr;
    std::vector<std::vector<uint8_t>> h_fac;     // [layer]
    std::vector<uint16_t> h_fmask;               // [layer][chunk][4096]
    uint16_t enemies[16] = {0};                  // enemies[f] = factions at war with f (G_GetEnemyFactions)
    bool faction_enabled = false;
    int req_faction = 0xF;                       // faction of the path request being planned (n_request_path's faction_id)
    uint8_t  *d_unit = nullptr;      // [layer][chunk] 1 if every passable cost in the chunk == 1
    std::vector<uint8_t> h_unit;     // host mirror
    CUtensorMap tmap_cost, tmap_blk; // rank-3 {x, y, layer}, box 64x64x1
    void *d_stage = nullptr; size_t stage_bytes = 0;   // upload staging
    unsigned long long *d_los_trace = nullptr; size_t los_trace_cap = 0, los_trace_n = 0; bool los_trace_on = false;
    int los_variant = 1;
    void *d_los_sched = nullptr; size_t los_sched_bytes = 0;   // LOS scheduler: work counter + per-request done flags
    cudaEvent_t ev_los_sched = nullptr;                        // recorded after every LOS launch: the next one waits for it
    // host mirrors (chunk-blocked, [layer][chunk][64][64]) for the host-side planner
    std::vector<uint8_t>  h_cost;
    std::vector<uint16_t> h_blk, h_liid;
    struct portal_t { int16_t chunk_r, chunk_c, r0, c0, r1, c1; int32_t conn_chunk, conn_idx; };
    std::vector<std::vector<std::vector<portal_t>>> portals;   // [layer][chunk][idx]
    // host-side derived state, owned by the context (no process-global tables: contexts on different threads are independent)
    void *route_state = nullptr;                               // pfnav_route.cu: std::vector<pfnav_route_layer>
    void *route_dev_state = nullptr;                           // pfnav_route.cu: device copies of the routing tables (k_portal_graph_path)
    std::set<std::pair<int, int>> dirty, fdirty;               // pfnav_blockers.cu: (layer, chunk) occupancy / faction mask changed
    void *blk_state = nullptr;                                 // pfnav_blockers.cu: queued blocker ops + device-side refcount state

    // ---- field pool ----
    int pool_ndests = 0, pool_max = 0, pool_used = 0;
    int32_t *d_pool_slot = nullptr;   // [ndests][chunks] -> slot or -1
    std::vector<int32_t> h_pool_slot;
    std::vector<uint8_t> h_pool_has;
    std::vector<pfnav_field_req> h_pool_req;   // [slot]: the last request applied to the slot (== flow_field::target, field.c:2076)
    std::vector<uint64_t> h_pool_ffid;     // [ndests][chunks]: ff_id currently mapped (dest, chunk) -> field, 0 = none
    void *d_plan_buf = nullptr; size_t plan_buf_bytes = 0;    // request staging for pfnav_pool_request_goal
    // The plan of the last goal batch stays resident on the device: re-requesting the same goals on an
    // unchanged map (map_epoch) relaunches the kernels without re-planning or re-uploading anything.
    uint64_t map_epoch = 0;
    struct goal_batch_t {
        bool valid = false; uint64_t epoch = 0; int layer = 0; std::vector<int32_t> dests, targets;
        int nf = 0, nl = 0; std::vector<int32_t> fwave_off; size_t b_fr = 0, b_fs = 0, b_lr = 0, b_ls = 0;
    } goal_batch;
    uint8_t *d_pool_flow = nullptr;   // [max][4096]
    uint8_t *d_pool_los = nullptr;    // [max][4096]
    // LRU eviction (fieldcache.c:59-71 keeps CONFIG_*_CACHE_SZ entries per LRU cache): every slot remembers the last
    // tick that read it (device side, written by the desired-velocity kernels) or requested it (host side)
    uint32_t *d_pool_touch = nullptr;             // [max] last tick number a work item read the slot
    std::vector<uint32_t> h_slot_touch;           // [max] last tick number a request named the slot
    std::vector<int64_t>  h_slot_owner;           // [max] dest * chunks + chunk, or -1 (free)
    std::vector<int32_t>  pool_free;              // evicted slots, reused before pool_used grows
    uint32_t tick_no = 1;                         // advanced by pfnav_agents_tick
    uint64_t pool_evictions = 0;
    // TARGET_ENEMIES / TARGET_ENTITY destinations (pfnav_pool_request_entity_fields): what their repairs start from
    struct aux_target { int kind = -1, layer = 0, ref_layer = 0; std::vector<pfnav_footprint> ents; };
    std::vector<aux_target> aux;

    // ---- agents ----
    size_t n_agents = 0, cap_agents = 0, n_flocks = 0, cap_flocks = 0;
    int hz = 20;
    pfnav_agent *d_agents = nullptr;      // AoS input as uploaded
    pf_record   *d_records = nullptr;     // 24-B records (uid order)
    pfnav_flock *d_flocks = nullptr;
    uint32_t *d_flock_start = nullptr;    // [nflocks+1] offsets into d_flock_members
    uint32_t *d_flock_members = nullptr;  // agent ids grouped by flock, ascending uid
    int32_t  *d_flock_of = nullptr;       // [n_agents] flock id column (all-gathered with the records)
    pfnav_formation_in  *d_formation = nullptr; size_t cap_formation = 0;   // optional formation inputs, uid order
    pfnav_movestate_ext *d_ms_ext = nullptr;    size_t cap_ms_ext = 0;      // optional: the movestate beyond point seeking
    void *d_enter = nullptr; size_t cap_enter = 0; void *d_ttiles = nullptr; size_t cap_ttiles = 0;   // ENTER_ENTITY_RANGE scratch
    uint32_t *d_facts = nullptr;          // {max radius bits, any garrisoned}
    // multi-GPU: this context owns the entity index range [shard_lo, shard_hi) of the population (pfnav_mgpu.cu)
    size_t shard_lo = 0, shard_hi = 0;
    void *mgpu = nullptr;                 // pf_mgpu: NCCL communicator or in-process group membership
    bool members_stale = false;           // flock member lists must be rebuilt at the next gather
    float2   *d_cohesion = nullptr;       // per-agent cohesion force (pre-pass)
    float2   *d_member_pos = nullptr; size_t cap_member_pos = 0;   // positions in flock-member order
    void *d_prep = nullptr; size_t cap_prep = 0;                   // phase-A results of the two-phase velocity update
    bool two_phase = true, two_phase_force = false;
    int phase_a_ctas_per_sm = 0;
    // spatial index (bitmap_grid.h equivalent)
    int grid_w = 0, grid_h = 0; int32_t origin_x = 0, origin_y = 0;
    uint32_t *d_cell_count = nullptr, *d_cell_start = nullptr, *d_cell_fill = nullptr;
    size_t cap_cells = 0;
    int32_t  *d_sorted_ix = nullptr, *d_sorted_iy = nullptr;
    uint32_t *d_sorted_id = nullptr;
    int32_t  *d_sorted_flock = nullptr;                  // flock id of every index entry (windowed cohesion)
    uint32_t *d_coh_fallback = nullptr; size_t cap_coh = 0;
    void *d_scan_part = nullptr;                              // block sums of the cell-count scan
    void *d_cp_stats = nullptr;                               // ClearPath event counters (pfnav_agents_clearpath_stats)
    void *d_coh_part = nullptr; size_t cap_coh_part = 0;      // windowed cohesion: partial sums of split runs, float4 [nsplit][n]
    std::vector<uint32_t> h_flock_start;
    int cohesion_mode = 0;
    void *d_scan_tmp = nullptr; size_t scan_tmp_bytes = 0;
    // work list + outputs
    size_t n_work = 0, cap_work = 0;
    uint32_t *d_work = nullptr;
    float2 *d_vel_out = nullptr, *d_vpref_out = nullptr, *d_vdes_out = nullptr;
    uint8_t *d_los_out = nullptr;
    uint32_t *d_work_count = nullptr;
    // state update (entity_compute_update)
    pfnav_movestate *d_movestate = nullptr; size_t cap_movestate = 0; bool movestate_set = false;
    pfnav_patch *d_patches = nullptr; size_t cap_patches = 0;
    void *d_arrival = nullptr; size_t arrival_bytes = 0;     // [nflocks][nlayers] pf_arrival_dev + tile lists
    uint64_t arrival_epoch = ~0ull; bool arrival_valid = false;
    std::vector<uint8_t> flock_layer_used;                   // [nflocks][12]
    std::vector<pfnav_flock> h_flocks;
    float max_radius = 0.0f;
    bool has_unsupported_state = false;
    bool any_garrisoned = false;                             // G_Pos_EntsInCircleFrom filters + reorders (position.c:100)
    uint32_t *d_nb_scratch = nullptr; size_t nb_scratch_warps = 0;
    cudaEvent_t update_done = nullptr;
    cudaStream_t tick_stream = nullptr;
    cudaEvent_t tick_done = nullptr;
    // LOS chains of a goal batch run on their own stream so that the parts of the tick that do not read
    // fields (all-gather, position index, cohesion) overlap the latency-bound LOS dependency chain
    cudaStream_t field_stream = nullptr, flow_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_los = nullptr, ev_flow = nullptr;
    // measured durations of the previous LOS batch and the previous velocity update steer the two-phase choice
    cudaEvent_t ev_vel0 = nullptr, ev_vel1 = nullptr;
    bool vel_timed = false; float last_los_ms = -1.0f, last_vel_ms = -1.0f;
    bool los_inflight = false;

    // ---- optional per-kernel timing (pfnav_profile_enable) ----
    bool profiling = false;
    struct prof_rec { int slot; cudaEvent_t a, b; };
    std::vector<prof_rec> prof_pending;
};

// Every entry point that takes a `void *stream`: NULL names the context's own non-blocking stream
// (never the legacy default stream, which would not be ordered against it).
static inline cudaStream_t pf_stream(pfnav_ctx *ctx, void *stream)
{
    return stream ? (cudaStream_t)stream : ctx->tick_stream;
}

// order `st` after the LOS work forked onto field_stream (no-op when none was forked)
static inline cudaError_t pf_fields_join(pfnav_ctx *ctx, cudaStream_t st)
{
    if (!ctx->los_inflight) return cudaSuccess;
    cudaError_t e = cudaStreamWaitEvent(st, ctx->ev_flow, 0);
    if (e != cudaSuccess) return e;
    return cudaStreamWaitEvent(st, ctx->ev_los, 0);
}
// host-side wait for the forked LOS work (setup-time entry points that use blocking copies)
static inline cudaError_t pf_fields_sync(pfnav_ctx *ctx)
{
    if (!ctx->los_inflight) return cudaSuccess;
    cudaError_t e = cudaEventSynchronize(ctx->ev_flow);
    if (e == cudaSuccess) e = cudaEventSynchronize(ctx->ev_los);
    if (e == cudaSuccess) ctx->los_inflight = false;
    return e;
}

// per (flock, layer) constants of arrived() (movement.c:2170), see pfnav_route.cu
#define PF_ARRIVAL_MC_MAX 256           /* FIELD_RES_R*2 + FIELD_RES_C*2 (nav.c:4721) */
struct pf_arrival_consts {
    int32_t nearest_ok; float nearest[2]; int32_t mc_n;
    float mc[PF_ARRIVAL_MC_MAX][2];
};
int pfnav_arrival_consts(pfnav_ctx *ctx, int layer, float tx, float tz, pf_arrival_consts *out);

// seeds of the flow-field repair chain (pfnav_route.cu), 64 rows x 64 bits
int pfnav_repair_seeds(pfnav_ctx *ctx, const pfnav_field_req &q, int kind, int arg, uint64_t *mask);

int pfnav_flow_repair_pool(pfnav_ctx *ctx, const pfnav_field_req *targets, const int32_t *kinds, const int32_t *args,
                           const int32_t *slots, size_t n);

int pfnav_fmask_push_chunk(pfnav_ctx *ctx, int layer, int chunk);
int pfnav_blockers_flush(pfnav_ctx *ctx);                        // pfnav_blockers.cu: apply the queued blocker operations now
int pfnav_blockers_factions_uploaded(pfnav_ctx *ctx, int layer); // pfnav_blockers.cu: host faction counts of a layer were replaced

int pfnav_aux_chunk_seeds(pfnav_ctx *ctx, int dest, int chunk_r, int chunk_c, std::vector<int> &out);     // pfnav_region.cu

// pool slots with LRU eviction (pfnav_plan.cu)
int pf_pool_reserve(pfnav_ctx *ctx, const size_t *keys, size_t n, int32_t *slots_out, bool *out_evicted);

enum { PF_PROF_FLOW = 0, PF_PROF_LOS = 1, PF_PROF_INDEX = 2, PF_PROF_VDES = 3, PF_PROF_COHESION = 4,
       PF_PROF_VELOCITY = 5, PF_PROF_UPDATE = 6, PF_PROF_APPLY = 7, PF_PROF_SLOTS = 8 };

// RAII helper: brackets the launches of one kernel group with events when profiling is on
struct pf_prof_scope {
    pfnav_ctx *ctx; cudaStream_t st; int slot; cudaEvent_t a = nullptr, b = nullptr;
    pf_prof_scope(pfnav_ctx *c, cudaStream_t s, int sl) : ctx(c), st(s), slot(sl)
    {
        if (!ctx->profiling) return;
        cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a, st);
    }
    ~pf_prof_scope()
    {
        if (!a) return;
        cudaEventRecord(b, st);
        ctx->prof_pending.push_back({slot, a, b});
    }
};

// ---- pfnav_fields.cu ----
int pfnav_fields_init(pfnav_ctx *ctx);
void pfnav_fields_free(pfnav_ctx *ctx);

int pfnav_flow_launch(pfnav_ctx *ctx, const pfnav_field_req *d_reqs, size_t n, uint8_t *d_inout_fields,
                      const int32_t *d_out_slot, void *stream);
int pfnav_los_launch(pfnav_ctx *ctx, const pfnav_los_req *d_reqs, size_t n, uint8_t *d_out_fields,
                     const int32_t *d_out_slot, int n_waves, const int32_t *h_wave_offsets, void *stream);

// ---- pfnav_agents.cu ----
void pfnav_agents_free(pfnav_ctx *ctx);
int pfnav_agents_init(pfnav_ctx *ctx);
int pfnav_agents_finish_snapshot(pfnav_ctx *ctx, cudaStream_t st, bool members);

// ---- device helpers shared by kernels ----
__device__ __forceinline__ uint32_t pf_lane() { return threadIdx.x & 31; }

// entity footprint -> tiles (pfnav_blockers.cu), for the TARGET_ENTITY / TARGET_ENEMIES frontiers
int pfnav_footprint_tiles(const pfnav_ctx *ctx, const pfnav_footprint *e, int rings, int32_t *out_rc);
