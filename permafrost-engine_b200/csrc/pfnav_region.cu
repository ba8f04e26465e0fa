// pfnav_region.cu -- region ("cell arrival" / "group arrival") fields, sm_100a.
//
// Reference: N_CellArrivalFieldCreate (navigation/field.c:2445), N_GroupArrivalFieldCreate (field.c:2525),
// N_CellArrivalFieldUpdateToNearestPathable (field.c:2603) and their helpers field_build_integration_region
// (:587), field_build_integration_nonpass_region (:678), field_passable_frontier (:1441), clamped_region
// (:1892), field_build_flow_unaligned (:804), field_flow_dir (:355), set_flow_cell (:790).
//
// One CTA per field. The dim x dim window (dim = 96 for formations) is gathered from the row-major layer
// images into shared memory once per pass: a cost byte and a flag byte per cell plus a 32-bit integration
// value (6 bytes per cell, 54 KB at dim 96, four CTAs per SM). The reference's Dijkstra has non-negative
// integer edge weights (cost_base of the tile entered, every sum < 2^24 so its float arithmetic is exact),
// so its result is the unique shortest-distance fixed point: the kernel reaches it by in-place relaxation
// sweeps (each thread owns a run of consecutive cells and alternates sweep direction) until no cell changes.
//
// The fix-up's flood (the passable rim of the blocked island around `start`) is a reachability fixed point as
// well -- except when the map-clamped region has fewer rows than columns: the reference indexes its visited
// array with stride region.r (visited_idx, field.c:1431), distinct tiles alias, and the outcome depends on
// the breadth-first order. That case (regions hanging over the top / bottom map edge only) is replayed
// literally by one thread, with the queue and visited array overlaid on the integration buffer.
#include <algorithm>
#include <cmath>
#include <cstring>
#include "pfnav_internal.cuh"

namespace {

#include "pfnav_region_kernel.cuh"
}   // namespace

static int region_launch(pfnav_ctx *ctx, int dim, const pfnav_region_req *d_reqs, size_t n, const int32_t *d_seeds,
                         const int32_t *d_overlay, uint8_t *d_fields, cudaStream_t st, int chunk_out = 0)
{
    RegionGrids g;
    g.cost = ctx->d_cost; g.blk = ctx->d_blk; g.fmask = ctx->d_fmask; g.W64 = ctx->W64; g.H64 = ctx->H64;
    const size_t smem = (size_t)dim * dim * 6;
    PF_CUDA(cudaFuncSetAttribute(k_region_fields, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 PFNAV_REGION_DIM_MAX * PFNAV_REGION_DIM_MAX * 6));
    const int per_sm = std::max(1, std::min(4, (int)((size_t)(227 * 1024) / (smem + 1024))));
    const int grid = (int)std::min<size_t>(n, (size_t)ctx->sm_count * per_sm);
    k_region_fields<<<grid, RG_THREADS, smem, st>>>(g, dim, d_reqs, (int)n, d_seeds, d_overlay, d_fields, chunk_out);
    ctx->launches++;
    PF_CUDA(cudaGetLastError());
    return PFNAV_OK;
}

extern "C" int pfnav_region_fields_dev(pfnav_ctx *ctx, int dim, const pfnav_region_req *d_reqs, size_t n,
                                       const int32_t *d_seeds_rc, const int32_t *d_overlay_rc, uint8_t *d_inout_fields,
                                       void *stream)
{
    PF_ARG(ctx && ctx->d_cost, "map not created");
    PF_NEED_DEVICE(ctx);
    if (n == 0) return PFNAV_OK;
    PF_ARG(dim >= 2 && dim <= PFNAV_REGION_DIM_MAX && dim % 2 == 0, "dim must be even and <= PFNAV_REGION_DIM_MAX");
    PF_ARG(d_reqs && d_inout_fields, "null buffer");
    PF_ARG(n < (1u << 30), "n");
    PF_CUDA(cudaSetDevice(ctx->device));
    return region_launch(ctx, dim, d_reqs, n, d_seeds_rc, d_overlay_rc, d_inout_fields, pf_stream(ctx, stream));
}

// the reference asserts these (field.c:2498-2499, 1455, 1487); here they are argument errors
static int validate_region_reqs(const pfnav_ctx *ctx, int dim, const pfnav_region_req *reqs, size_t n, const int32_t *seeds,
                                size_t nseeds, size_t noverlay)
{
    const int half = dim / 2;
    for (size_t i = 0; i < n; i++) {
        const pfnav_region_req &q = reqs[i];
        PF_ARG(q.layer >= 0 && q.layer < ctx->nlayers, "region request: layer");
        PF_ARG((q.flags & (PFNAV_REGION_CREATE | PFNAV_REGION_FIXUP)) != 0, "region request: neither CREATE nor FIXUP");
        PF_ARG(q.center_r >= 0 && q.center_r < ctx->H64 && q.center_c >= 0 && q.center_c < ctx->W64, "region request: center outside the map");
        PF_ARG(q.overlay_n >= 0 && q.overlay_off >= 0 && (size_t)q.overlay_off + q.overlay_n <= noverlay, "region request: overlay range");
        if (q.flags & PFNAV_REGION_CREATE) {
            PF_ARG(q.seed_n >= 0 && q.seed_off >= 0 && (size_t)q.seed_off + q.seed_n <= nseeds, "region request: seed range");
            for (int k = 0; k < q.seed_n; k++) {
                const int32_t *s = seeds + 2 * ((size_t)q.seed_off + k);
                PF_ARG(s[0] >= 0 && s[0] < ctx->H64 && s[1] >= 0 && s[1] < ctx->W64, "region request: seed tile outside the map");
            }
            if (q.flags & PFNAV_REGION_CELL) {
                PF_ARG(q.seed_n == 1, "region request: PFNAV_REGION_CELL takes exactly one seed");
                const int32_t *s = seeds + 2 * (size_t)q.seed_off;
                PF_ARG(s[0] >= q.center_r - half && s[1] >= q.center_c - half, "region request: cell tile before the region base");
            }
        }
        if (q.flags & PFNAV_REGION_FIXUP) {
            const int cb_r = std::max(q.center_r - half, 0), cb_c = std::max(q.center_c - half, 0);
            const int ce_r = q.center_r + half < ctx->H64 ? q.center_r + half : ctx->H64 - 1;
            const int ce_c = q.center_c + half < ctx->W64 ? q.center_c + half : ctx->W64 - 1;
            PF_ARG(q.start_r >= cb_r && q.start_r < ce_r && q.start_c >= cb_c && q.start_c < ce_c,
                   "region request: fix-up start outside the clamped region");
        }
    }
    return PFNAV_OK;
}

extern "C" int pfnav_region_fields(pfnav_ctx *ctx, int dim, const pfnav_region_req *reqs, size_t n, const int32_t *seeds_rc,
                                   size_t nseeds, const int32_t *overlay_rc, size_t noverlay, uint8_t *inout_fields)
{
    PF_ARG(ctx && ctx->d_cost, "map not created");
    PF_NEED_DEVICE(ctx);
    if (n == 0) return PFNAV_OK;
    PF_ARG(dim >= 2 && dim <= PFNAV_REGION_DIM_MAX && dim % 2 == 0, "dim must be even and <= PFNAV_REGION_DIM_MAX");
    PF_ARG(reqs && inout_fields, "null buffer");
    PF_ARG((nseeds == 0 || seeds_rc) && (noverlay == 0 || overlay_rc), "null seed / overlay list");
    PF_ARG(n < (1u << 30), "n");
    int rc = validate_region_reqs(ctx, dim, reqs, n, seeds_rc, nseeds, noverlay);
    if (rc) return rc;
    PF_CUDA(cudaSetDevice(ctx->device));
    const size_t fbytes = (size_t)dim * dim / 2;
    bool need_in = false;
    for (size_t i = 0; i < n; i++) need_in |= !(reqs[i].flags & PFNAV_REGION_CREATE);
    const size_t b_req = n * sizeof(pfnav_region_req), b_seed = std::max<size_t>(nseeds, 1) * 8, b_ov = std::max<size_t>(noverlay, 1) * 8;
    uint8_t *d_buf = nullptr, *d_fields = nullptr;
    PF_CUDA(cudaMalloc(&d_buf, b_req + b_seed + b_ov));
    if (cudaMalloc(&d_fields, n * fbytes) != cudaSuccess) { cudaFree(d_buf); pfnav_set_error("cudaMalloc region fields"); return PFNAV_ERR_NOMEM; }
    pfnav_region_req *d_reqs = reinterpret_cast<pfnav_region_req *>(d_buf);
    int32_t *d_seeds = reinterpret_cast<int32_t *>(d_buf + b_req), *d_ov = reinterpret_cast<int32_t *>(d_buf + b_req + b_seed);
    cudaStream_t st = ctx->tick_stream;
    cudaError_t e = cudaMemcpyAsync(d_reqs, reqs, b_req, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && nseeds) e = cudaMemcpyAsync(d_seeds, seeds_rc, nseeds * 8, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && noverlay) e = cudaMemcpyAsync(d_ov, overlay_rc, noverlay * 8, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && need_in) e = cudaMemcpyAsync(d_fields, inout_fields, n * fbytes, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        rc = region_launch(ctx, dim, d_reqs, n, d_seeds, d_ov, d_fields, st);
        if (rc == 0) e = cudaMemcpyAsync(inout_fields, d_fields, n * fbytes, cudaMemcpyDeviceToHost, st);
    }
    cudaError_t e2 = cudaStreamSynchronize(st);
    cudaFree(d_buf); cudaFree(d_fields);
    if (rc) return rc;
    if (e != cudaSuccess || e2 != cudaSuccess) {
        pfnav_set_error("pfnav_region_fields: %s", cudaGetErrorString(e != cudaSuccess ? e : e2));
        return PFNAV_ERR_CUDA;
    }
    return PFNAV_OK;
}

// M_Tile_DescForPoint2D with the nav resolution (tile.c:547) -> absolute (r, c)
static bool region_tile_for_point(const pfnav_ctx *ctx, float px, float pz, int32_t *out_rc)
{
    const float width = (float)(ctx->chunk_w * 256), height = (float)(ctx->chunk_h * 256);
    if (px > ctx->map_x || px < ctx->map_x - width) return false;
    if (pz < ctx->map_z || pz > ctx->map_z + height) return false;
    int chunk_r = (int)(fabs(ctx->map_z - pz) / 256.0f), chunk_c = (int)(fabs(ctx->map_x - px) / 256.0f);
    chunk_r = std::min(std::max(chunk_r, 0), ctx->chunk_h - 1);
    chunk_c = std::min(std::max(chunk_c, 0), ctx->chunk_w - 1);
    const float bx = ctx->map_x - (chunk_c * 256.0f), bz = ctx->map_z + (chunk_r * 256.0f);
    const int tile_r = (int)(fabs(bz - pz) / 4), tile_c = (int)(fabs(bx - px) / 4);
    out_rc[0] = chunk_r * 64 + std::min(std::max(tile_r, 0), 63);
    out_rc[1] = chunk_c * 64 + std::min(std::max(tile_c, 0), 63);
    return true;
}

extern "C" int pfnav_group_arrival_field(pfnav_ctx *ctx, int layer, int dim, uint16_t enemies, const float *targets_xz,
                                         size_t ntargets, const float *center_xz, const int32_t *overlay_rc, size_t noverlay,
                                         uint8_t *out_field)
{
    PF_ARG(ctx && ctx->d_cost, "map not created");
    PF_NEED_DEVICE(ctx);
    PF_ARG(dim >= 2 && dim <= PFNAV_REGION_DIM_MAX && dim % 2 == 0, "dim must be even and <= PFNAV_REGION_DIM_MAX");
    PF_ARG(center_xz && out_field && (ntargets == 0 || targets_xz), "null buffer");
    PF_ARG(ntargets < (1u << 28), "ntargets");
    int32_t c[2];
    if (!region_tile_for_point(ctx, center_xz[0], center_xz[1], c)) {        // field.c:2552-2556
        memset(out_field, 0, (size_t)dim * dim / 2);
        return PFNAV_OK;
    }
    std::vector<int32_t> seeds;
    seeds.reserve(ntargets * 2);
    for (size_t i = 0; i < ntargets; i++) {
        int32_t t[2];
        if (region_tile_for_point(ctx, targets_xz[2 * i], targets_xz[2 * i + 1], t)) { seeds.push_back(t[0]); seeds.push_back(t[1]); }
    }
    pfnav_region_req q = {};
    q.layer = layer; q.center_r = c[0]; q.center_c = c[1];
    q.seed_off = 0; q.seed_n = (int32_t)(seeds.size() / 2);
    q.overlay_off = 0; q.overlay_n = (int32_t)noverlay;
    q.enemies = enemies; q.flags = PFNAV_REGION_CREATE;
    return pfnav_region_fields(ctx, dim, &q, 1, seeds.data(), seeds.size() / 2, overlay_rc, noverlay, out_field);
}

// ------------------------------------------------------------------------------------------
// TARGET_ZONE chunk fields: the group arrival fields (N_RequestAsyncGroupArrivalField nav.c:3921 ->
// N_FlowFieldUpdate field.c:2050 -> field_update_zone :1810) and their per-entity consumer
// N_DesiredGroupArrivalVelocity (nav.c:3561).
// ------------------------------------------------------------------------------------------
namespace {

// lib/public/pqueue.h:109-208 (1-indexed binary min-heap, hole-based sift; ties keep the reference's order)
struct zone_heap {
    struct node { float prio; int r, c; };
    std::vector<node> a{1};
    size_t size() const { return a.size() - 1; }
    void clear() { a.resize(1); }
    void push(float prio, int r, int c)
    {
        a.push_back({});
        size_t curr = a.size() - 1, parent = curr / 2;
        while (curr > 1 && a[parent].prio > prio) { a[curr] = a[parent]; curr = parent; parent /= 2; }
        a[curr] = {prio, r, c};
    }
    void pop(int *r, int *c)
    {
        *r = a[1].r; *c = a[1].c;
        const size_t n = a.size() - 2;      // size after the pop; a[n + 1], the last element, is the one being sifted
        a[1] = a[n + 1];
        size_t root = 1;
        while (root != n + 1) {             // _pq_balance: `target` starts at the sifted element's own slot
            size_t target = n + 1;
            const size_t l = root * 2, rr = l + 1;
            if (l <= n && a[l].prio < a[target].prio) target = l;
            if (rr <= n && a[rr].prio < a[target].prio) target = rr;
            a[root] = a[target];
            root = target;
        }
        a.pop_back();
    }
};

struct zone_map {
    const uint8_t *cost; const uint16_t *blk; int chunk_w, chunk_h;
    bool exists(int r, int c) const { return r >= 0 && r < chunk_h * 64 && c >= 0 && c < chunk_w * 64; }
    bool passable(int r, int c) const       // field_tile_passable (field.c:117)
    {
        const size_t off = ((size_t)(r / 64) * chunk_w + c / 64) * 4096 + (r % 64) * 64 + (c % 64);
        return cost[off] != 0xFF && blk[off] == 0;
    }
};

// field_zone_initial_frontier (field.c:1683)
void zone_initial_frontier(const zone_map &m, int centre_r, int centre_c, int base_r, int base_c, int dim, size_t budget,
                           std::vector<int32_t> &out)
{
    static const int er[8] = {0, 0, -1, 1, -1, -1, 1, 1}, ec[8] = {-1, 1, 0, 0, -1, 1, -1, 1};
    const int cdr = centre_r - base_r, cdc = centre_c - base_c;
    if (cdr < 0 || cdr >= dim || cdc < 0 || cdc >= dim) return;
    std::vector<uint8_t> visited((size_t)dim * dim, 0);
    zone_heap frontier;
    visited[cdr * dim + cdc] = 1;
    frontier.push(0.0f, centre_r, centre_c);
    int start_r = 0, start_c = 0; bool have_start = false;
    while (frontier.size() > 0) {
        int r, c; frontier.pop(&r, &c);
        if (m.passable(r, c)) { start_r = r; start_c = c; have_start = true; break; }
        for (int e = 0; e < 8; e++) {
            const int nr = r + er[e], nc = c + ec[e];
            if (!m.exists(nr, nc)) continue;
            const int dr = nr - base_r, dc = nc - base_c;
            if (dr < 0 || dr >= dim || dc < 0 || dc >= dim) continue;
            if (visited[dr * dim + dc]) continue;
            visited[dr * dim + dc] = 1;
            const int ndr = nr - centre_r, ndc = nc - centre_c;
            frontier.push((float)(ndr * ndr + ndc * ndc), nr, nc);
        }
    }
    if (!have_start) return;
    frontier.clear();
    std::fill(visited.begin(), visited.end(), 0);
    visited[(start_r - base_r) * dim + (start_c - base_c)] = 1;
    frontier.push(0.0f, start_r, start_c);
    size_t ret = 0;
    while (frontier.size() > 0 && ret < budget) {
        int r, c; frontier.pop(&r, &c);
        if (m.passable(r, c)) { out.push_back(r); out.push_back(c); ret++; }
        for (int e = 0; e < 8; e++) {
            const int nr = r + er[e], nc = c + ec[e];
            if (!m.exists(nr, nc)) continue;
            const int dr = nr - base_r, dc = nc - base_c;
            if (dr < 0 || dr >= dim || dc < 0 || dc >= dim) continue;
            if (visited[dr * dim + dc]) continue;
            visited[dr * dim + dc] = 1;
            if (!m.passable(nr, nc)) continue;
            const int ndr = nr - centre_r, ndc = nc - centre_c;
            frontier.push((float)(ndr * ndr + ndc * ndc), nr, nc);
        }
    }
}

struct ZoneView {
    const int32_t *slot; const uint8_t *flow; const uint8_t *has;     // the field pool (pfnav_pool_*)
    int chunk_w, chunk_h; float map_x, map_z;
};

// M_Tile_DescForPoint2D (tile.c:547) -> absolute (r, c); divisions by powers of two are exact in float
__device__ __forceinline__ bool zone_tile_for_point(const ZoneView &z, float px, float pz, int &ar, int &ac)
{
    const float width = (float)(z.chunk_w * 256), height = (float)(z.chunk_h * 256);
    if (px > z.map_x || px < z.map_x - width) return false;
    if (pz < z.map_z || pz > z.map_z + height) return false;
    int chunk_r = (int)(fabsf(z.map_z - pz) / 256.0f), chunk_c = (int)(fabsf(z.map_x - px) / 256.0f);
    chunk_r = min(max(chunk_r, 0), z.chunk_h - 1);
    chunk_c = min(max(chunk_c, 0), z.chunk_w - 1);
    const float bx = z.map_x - (float)chunk_c * 256.0f, bz = z.map_z + (float)chunk_r * 256.0f;
    const int tile_r = (int)(fabsf(bz - pz) / 4.0f), tile_c = (int)(fabsf(bx - px) / 4.0f);
    ar = chunk_r * 64 + min(max(tile_r, 0), 63);
    ac = chunk_c * 64 + min(max(tile_c, 0), 63);
    return true;
}

// N_DesiredGroupArrivalVelocity (nav.c:3561): one thread per position
__global__ void k_group_arrival_velocity(ZoneView z, int dest, float cx, float cz, int radius, const float2 *__restrict__ pos, int n,
                                         float2 *__restrict__ out_vel, uint8_t *__restrict__ out_flags)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float2 v = make_float2(0.0f, 0.0f);
    uint8_t fl = 0;
    int ar, ac, cr, cc;
    const float2 p = pos[i];
    if (zone_tile_for_point(z, p.x, p.y, ar, ac) && zone_tile_for_point(z, cx, cz, cr, cc)) {
        const int chunk = (ar >> 6) * z.chunk_w + (ac >> 6);
        const int slot = z.slot[(size_t)dest * z.chunk_w * z.chunk_h + chunk];
        if (slot >= 0 && (z.has[slot] & 1)) {
            const uint32_t dir = z.flow[(size_t)slot * 4096 + (ar & 63) * 64 + (ac & 63)];
            // N_FlowDir (field.c:2429): x grows to the left
            const float d = (float)(1.0 / 1.4142135623730951);      // 1.0f / sqrt(2.0f): a double division rounded to float
            switch (dir) {
            case 1: v = make_float2(d, -d); break;
            case 2: v = make_float2(0.0f, -1.0f); break;
            case 3: v = make_float2(-d, -d); break;
            case 4: v = make_float2(1.0f, 0.0f); break;
            case 5: v = make_float2(-1.0f, 0.0f); break;
            case 6: v = make_float2(d, d); break;
            case 7: v = make_float2(0.0f, 1.0f); break;
            case 8: v = make_float2(-d, d); break;
            default: break;
            }
            fl = 1;
            if (dir == 0) {
                const int dr = ar - cr, dc = ac - cc;
                if (dr * dr + dc * dc <= radius * radius) fl |= 2;
            }
        }
    }
    out_vel[i] = v; out_flags[i] = fl;
}

}   // namespace

static int zone_dim(const pfnav_ctx *ctx, int *out_dim)
{
    // field_update_zone sizes its square buffer with rdim as the row stride (field.c:1821-1828): only the
    // cases where rows == columns are defined
    if (ctx->chunk_h > 1 && ctx->chunk_w > 1) { *out_dim = 128; return PFNAV_OK; }
    if (ctx->chunk_h == 1 && ctx->chunk_w == 1) { *out_dim = 64; return PFNAV_OK; }
    pfnav_set_error("zone fields need a map with more than one chunk row AND column (or a 1 x 1 map): the reference's padded "
                    "region is indexed with one stride for both");
    return PFNAV_ERR_ARG;
}

// Padded-chunk fields from per-chunk seed lists (seed_off[i] .. seed_off[i + 1] pairs of chunk i); d_out = n x 4096
// direction bytes. Shared by the zone, entity and enemies targets (field_update_zone / _entity / _enemies).
static int chunk_fields_launch(pfnav_ctx *ctx, int layer, int dim, const int32_t *chunks_rc, size_t n, const std::vector<int32_t> &seeds,
                               const std::vector<size_t> &seed_off, uint8_t *d_out, cudaStream_t st)
{
    std::vector<pfnav_region_req> reqs(n);
    for (size_t i = 0; i < n; i++) {
        pfnav_region_req q = {};
        q.layer = layer; q.center_r = chunks_rc[2 * i]; q.center_c = chunks_rc[2 * i + 1];
        q.seed_off = (int32_t)seed_off[i]; q.seed_n = (int32_t)(seed_off[i + 1] - seed_off[i]);
        q.flags = PFNAV_REGION_CREATE;
        reqs[i] = q;
    }
    const size_t b_req = n * sizeof(pfnav_region_req), b_seed = std::max<size_t>(seeds.size(), 2) * 4;
    uint8_t *d_buf = nullptr;
    PF_CUDA(cudaMalloc(&d_buf, b_req + b_seed + 8));
    int rc = PFNAV_OK;
    cudaError_t e = cudaMemcpyAsync(d_buf, reqs.data(), b_req, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && !seeds.empty()) e = cudaMemcpyAsync(d_buf + b_req, seeds.data(), seeds.size() * 4, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess)
        rc = region_launch(ctx, dim, reinterpret_cast<const pfnav_region_req *>(d_buf), n, reinterpret_cast<const int32_t *>(d_buf + b_req),
                           reinterpret_cast<const int32_t *>(d_buf + b_req + b_seed), d_out, st, 1);
    // the staging buffer is read by the kernel: free it once the stream has passed it
    cudaError_t e2 = cudaStreamSynchronize(st);
    cudaFree(d_buf);
    if (rc) return rc;
    if (e != cudaSuccess || e2 != cudaSuccess) {
        pfnav_set_error("chunk fields: %s", cudaGetErrorString(e != cudaSuccess ? e : e2));
        return PFNAV_ERR_CUDA;
    }
    return PFNAV_OK;
}

// the padded region of chunk (cr, cc): field.c:1566-1571
static inline void chunk_region_base(int dim, int cr, int cc, int *base_r, int *base_c)
{
    *base_r = (cr > 0 && dim > 64) ? cr * 64 - 32 : cr * 64;
    *base_c = (cc > 0 && dim > 64) ? cc * 64 - 32 : cc * 64;
}

// zone seeds of every requested chunk + the device launch
static int zone_fields_launch(pfnav_ctx *ctx, int layer, int centre_r, int centre_c, int radius, const int32_t *chunks_rc, size_t n,
                              uint8_t *d_out, cudaStream_t st)
{
    int dim = 0;
    int rc = zone_dim(ctx, &dim);
    if (rc) return rc;
    const size_t ltiles = (size_t)ctx->chunk_w * ctx->chunk_h * 4096;
    zone_map zm = { ctx->h_cost.data() + ltiles * layer, ctx->h_blk.data() + ltiles * layer, ctx->chunk_w, ctx->chunk_h };
    size_t budget = (size_t)(M_PI * radius * radius + 0.5);             // field.c:1843
    if (budget > (size_t)dim * dim) budget = (size_t)dim * dim;
    std::vector<int32_t> seeds;
    std::vector<size_t> off(n + 1, 0);
    for (size_t i = 0; i < n; i++) {
        const int cr = chunks_rc[2 * i], cc = chunks_rc[2 * i + 1];
        PF_ARG(cr >= 0 && cr < ctx->chunk_h && cc >= 0 && cc < ctx->chunk_w, "zone field: chunk");
        int base_r, base_c;
        chunk_region_base(dim, cr, cc, &base_r, &base_c);
        zone_initial_frontier(zm, centre_r, centre_c, base_r, base_c, dim, budget, seeds);
        off[i + 1] = seeds.size() / 2;
    }
    return chunk_fields_launch(ctx, layer, dim, chunks_rc, n, seeds, off, d_out, st);
}

extern "C" int pfnav_zone_fields(pfnav_ctx *ctx, int layer, int centre_r, int centre_c, int radius, const int32_t *chunks_rc,
                                 size_t n, uint8_t *out_fields)
{
    PF_ARG(ctx && ctx->d_cost, "map not created");
    PF_NEED_DEVICE(ctx);
    if (n == 0) return PFNAV_OK;
    PF_ARG(chunks_rc && out_fields, "null buffer");
    PF_ARG(layer >= 0 && layer < ctx->nlayers, "layer");
    PF_ARG(centre_r >= 0 && centre_r < ctx->H64 && centre_c >= 0 && centre_c < ctx->W64, "zone centre outside the map");
    PF_ARG(radius >= 0 && radius <= 0xFFFF, "radius");
    PF_ARG(n < (1u << 24), "n");
    PF_CUDA(cudaSetDevice(ctx->device));
    uint8_t *d_out = nullptr;
    PF_CUDA(cudaMalloc(&d_out, n * 4096));
    int rc = zone_fields_launch(ctx, layer, centre_r, centre_c, radius, chunks_rc, n, d_out, ctx->tick_stream);
    if (rc == 0 && cudaMemcpy(out_fields, d_out, n * 4096, cudaMemcpyDeviceToHost) != cudaSuccess) {
        pfnav_set_error("pfnav_zone_fields: copy back failed"); rc = PFNAV_ERR_CUDA;
    }
    cudaFree(d_out);
    return rc;
}

extern "C" int pfnav_pool_request_zone(pfnav_ctx *ctx, int dest, int layer, const float *centre_xz, int radius, void *stream,
                                       int *out_nfields)
{
    PF_ARG(ctx && ctx->d_cost, "map not created");
    PF_NEED_DEVICE(ctx);
    PF_ARG(ctx->d_pool_slot, "pool not created");
    PF_ARG(dest >= 0 && dest < ctx->pool_ndests, "dest");
    PF_ARG(layer >= 0 && layer < ctx->nlayers, "layer");
    PF_ARG(centre_xz, "centre");
    PF_ARG(radius >= 0 && radius <= 0xFFFF, "radius");
    if (out_nfields) *out_nfields = 0;
    int32_t c[2];
    if (!region_tile_for_point(ctx, centre_xz[0], centre_xz[1], c)) return PFNAV_OK;       // nav.c:3931
    PF_CUDA(cudaSetDevice(ctx->device));
    PF_CUDA(pf_fields_sync(ctx));
    cudaStream_t st = pf_stream(ctx, stream);
    // the chunks the footprint can reach (nav.c:3945-3951; C division truncates toward zero)
    const int reach = 2 * radius;
    auto clampi = [](int v, int lo, int hi) { return std::min(std::max(v, lo), hi); };
    const int min_cr = clampi((c[0] - reach) / 64, 0, ctx->chunk_h - 1), max_cr = clampi((c[0] + reach) / 64, 0, ctx->chunk_h - 1);
    const int min_cc = clampi((c[1] - reach) / 64, 0, ctx->chunk_w - 1), max_cc = clampi((c[1] + reach) / 64, 0, ctx->chunk_w - 1);
    std::vector<int32_t> chunks;
    for (int cr = min_cr; cr <= max_cr; cr++)
        for (int cc = min_cc; cc <= max_cc; cc++) { chunks.push_back(cr); chunks.push_back(cc); }
    const size_t n = chunks.size() / 2;
    const int nchunks = ctx->chunk_w * ctx->chunk_h;
    // pool slots first, so that a full pool fails before any work is queued
    std::vector<int32_t> slots(n);
    for (size_t i = 0; i < n; i++) {
        const size_t si = (size_t)dest * nchunks + chunks[2 * i] * ctx->chunk_w + chunks[2 * i + 1];
        int slot = ctx->h_pool_slot[si];
        if (slot < 0) {
            if (ctx->pool_used >= ctx->pool_max) { pfnav_set_error("pfnav_pool_request_zone: pool full (%d fields)", ctx->pool_max); return PFNAV_ERR_NOMEM; }
            slot = ctx->pool_used++;
            ctx->h_pool_slot[si] = slot;
            PF_CUDA(cudaMemcpyAsync(ctx->d_pool_slot + si, &ctx->h_pool_slot[si], sizeof(int32_t), cudaMemcpyHostToDevice, st));
        }
        slots[i] = slot;
    }
    uint8_t *d_tmp = nullptr;
    PF_CUDA(cudaMalloc(&d_tmp, n * 4096));
    int rc = zone_fields_launch(ctx, layer, c[0], c[1], radius, chunks.data(), n, d_tmp, st);
    uint8_t *d_has = ctx->d_pool_los + (size_t)ctx->pool_max * 4096;
    for (size_t i = 0; i < n && rc == 0; i++) {
        ctx->h_pool_has[slots[i]] |= 1;
        if (cudaMemcpyAsync(ctx->d_pool_flow + (size_t)slots[i] * 4096, d_tmp + i * 4096, 4096, cudaMemcpyDeviceToDevice, st) != cudaSuccess ||
            cudaMemcpyAsync(d_has + slots[i], &ctx->h_pool_has[slots[i]], 1, cudaMemcpyHostToDevice, st) != cudaSuccess) {
            pfnav_set_error("pfnav_pool_request_zone: pool copy failed"); rc = PFNAV_ERR_CUDA;
        }
    }
    cudaStreamSynchronize(st);
    cudaFree(d_tmp);
    ctx->goal_batch.valid = false;
    if (rc == 0 && out_nfields) *out_nfields = (int)n;
    return rc;
}

extern "C" int pfnav_group_arrival_velocity(pfnav_ctx *ctx, int dest, const float *centre_xz, int radius, const float *pos_xz,
                                            size_t n, float *out_vel, uint8_t *out_flags)
{
    PF_ARG(ctx && ctx->d_cost, "map not created");
    PF_NEED_DEVICE(ctx);
    PF_ARG(ctx->d_pool_slot, "pool not created");
    PF_ARG(dest >= 0 && dest < ctx->pool_ndests, "dest");
    if (n == 0) return PFNAV_OK;
    PF_ARG(centre_xz && pos_xz && out_vel && out_flags, "null buffer");
    PF_ARG(n < (1u << 30), "n");
    PF_ARG(radius >= 0 && radius <= 0xFFFF, "radius");
    PF_CUDA(cudaSetDevice(ctx->device));
    PF_CUDA(pf_fields_sync(ctx));
    uint8_t *d_buf = nullptr;
    PF_CUDA(cudaMalloc(&d_buf, n * 17));
    float2 *d_pos = reinterpret_cast<float2 *>(d_buf), *d_vel = d_pos + n;
    uint8_t *d_fl = d_buf + n * 16;
    cudaStream_t st = ctx->tick_stream;
    ZoneView z = { ctx->d_pool_slot, ctx->d_pool_flow, ctx->d_pool_los + (size_t)ctx->pool_max * 4096, ctx->chunk_w, ctx->chunk_h,
                   ctx->map_x, ctx->map_z };
    cudaError_t e = cudaMemcpyAsync(d_pos, pos_xz, n * 8, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        k_group_arrival_velocity<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(z, dest, centre_xz[0], centre_xz[1], radius, d_pos, (int)n, d_vel, d_fl);
        ctx->launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_vel, d_vel, n * 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_flags, d_fl, n, cudaMemcpyDeviceToHost, st);
    cudaError_t e2 = cudaStreamSynchronize(st);
    cudaFree(d_buf);
    if (e != cudaSuccess || e2 != cudaSuccess) {
        pfnav_set_error("pfnav_group_arrival_velocity: %s", cudaGetErrorString(e != cudaSuccess ? e : e2));
        return PFNAV_ERR_CUDA;
    }
    return PFNAV_OK;
}

// The zone's seed tiles alone (host structure code, no device needed): the first `budget` open tiles of the
// best-first flood. out_rc = (r, c) pairs, at most cap of them; *out_n = the full count.
extern "C" int pfnav_zone_seeds(pfnav_ctx *ctx, int layer, int chunk_r, int chunk_c, int centre_r, int centre_c, int radius,
                                int32_t *out_rc, size_t cap, size_t *out_n)
{
    PF_ARG(ctx && !ctx->h_cost.empty(), "map not created");
    PF_ARG(layer >= 0 && layer < ctx->nlayers, "layer");
    PF_ARG(chunk_r >= 0 && chunk_r < ctx->chunk_h && chunk_c >= 0 && chunk_c < ctx->chunk_w, "chunk");
    PF_ARG(centre_r >= 0 && centre_r < ctx->H64 && centre_c >= 0 && centre_c < ctx->W64, "zone centre outside the map");
    PF_ARG(radius >= 0 && radius <= 0xFFFF && out_n, "radius / out_n");
    int dim = 0;
    int rc = zone_dim(ctx, &dim);
    if (rc) return rc;
    const size_t ltiles = (size_t)ctx->chunk_w * ctx->chunk_h * 4096;
    zone_map zm = { ctx->h_cost.data() + ltiles * layer, ctx->h_blk.data() + ltiles * layer, ctx->chunk_w, ctx->chunk_h };
    size_t budget = (size_t)(M_PI * radius * radius + 0.5);
    if (budget > (size_t)dim * dim) budget = (size_t)dim * dim;
    const int base_r = (chunk_r > 0 && dim > 64) ? chunk_r * 64 - 32 : chunk_r * 64, base_c = (chunk_c > 0 && dim > 64) ? chunk_c * 64 - 32 : chunk_c * 64;
    std::vector<int32_t> seeds;
    zone_initial_frontier(zm, centre_r, centre_c, base_r, base_c, dim, budget, seeds);
    *out_n = seeds.size() / 2;
    if (out_rc) memcpy(out_rc, seeds.data(), std::min(cap * 2, seeds.size()) * sizeof(int32_t));
    return PFNAV_OK;
}


// ------------------------------------------------------------------------------------------
// TARGET_ENTITY / TARGET_ENEMIES chunk fields (field_update_entity field.c:1609, field_update_enemies :1540)
// ------------------------------------------------------------------------------------------
// BG_SCALE_F (lib/public/bitmap_grid.h:196): the position index compares scaled integers, bounds inclusive (:543)
static inline int32_t bg_scale_f(float x) { return (int32_t)lrintf(x * 256.0f); }

static int entity_seeds(const pfnav_ctx *ctx, int ref_layer, int kind, const pfnav_footprint *ents, size_t nents, int dim,
                        int chunk_r, int chunk_c, std::vector<int32_t> &out)
{
    int base_r, base_c;
    chunk_region_base(dim, chunk_r, chunk_c, &base_r, &base_c);
    int32_t tiles[512 * 2];
    if (kind == PFNAV_TARGET_ENTITY) {
        // field_entity_initial_frontier (field.c:1317): note `layer == GROUND_3X3` for the first ring
        const int rings = (ref_layer == 1) + (ref_layer >= 2) + (ref_layer >= 3);
        const int n = pfnav_footprint_tiles(ctx, &ents[0], rings, tiles);
        if (n < 0) { pfnav_set_error("entity field: a building corner lies outside the map"); return PFNAV_ERR_ARG; }
        for (int i = 0; i < n; i++) {
            const int dr = tiles[2 * i] - base_r, dc = tiles[2 * i + 1] - base_c;
            if (dr < 0 || dr >= dim || dc < 0 || dc >= dim) continue;
            out.push_back(tiles[2 * i]); out.push_back(tiles[2 * i + 1]);
        }
        return PFNAV_OK;
    }
    // field_enemies_initial_frontier (field.c:1209)
    const int rings = (ref_layer >= 1) + (ref_layer >= 2) + (ref_layer >= 3);
    // field_chunk_bounds (field.c:863) and the search rectangle (:1229-1243), in the reference's float arithmetic
    const int x_offset = -(chunk_c * 256), z_offset = chunk_r * 256;
    const float x_max = ctx->map_x + (float)x_offset, x_min = x_max - 256.0f;
    const float z_min = ctx->map_z + (float)z_offset, z_max = z_min + 256.0f;
    const float xlen = x_max - x_min, zlen = z_max - z_min;
    const int32_t imnx = bg_scale_f(x_min - xlen / 2.0f - 16.0f), imxx = bg_scale_f(x_max + xlen / 2.0f + 16.0f);
    const int32_t imnz = bg_scale_f(z_min - zlen / 2.0f - 16.0f), imxz = bg_scale_f(z_max + zlen / 2.0f + 16.0f);
    std::vector<uint8_t> has((size_t)dim * dim, 0);
    for (size_t k = 0; k < nents; k++) {
        const int32_t ix = bg_scale_f(ents[k].x), iz = bg_scale_f(ents[k].z);
        if (ix < imnx || ix > imxx || iz < imnz || iz > imxz) continue;
        const int n = pfnav_footprint_tiles(ctx, &ents[k], rings, tiles);
        if (n < 0) { pfnav_set_error("enemies field: a building corner lies outside the map"); return PFNAV_ERR_ARG; }
        for (int i = 0; i < n; i++) {
            const int dr = tiles[2 * i] - base_r, dc = tiles[2 * i + 1] - base_c;
            if (dr < 0 || dr >= dim || dc < 0 || dc >= dim) continue;
            has[(size_t)dr * dim + dc] = 1;
        }
    }
    for (int r = 0; r < dim; r++)
        for (int c = 0; c < dim; c++)
            if (has[(size_t)r * dim + c]) { out.push_back(base_r + r); out.push_back(base_c + c); }
    return PFNAV_OK;
}

static int entity_args_ok(const pfnav_ctx *ctx, int ref_layer, int kind, const pfnav_footprint *ents, size_t nents)
{
    PF_ARG(ref_layer >= 0 && ref_layer < PFNAV_NAV_LAYER_MAX, "ref_layer");
    PF_ARG(kind == PFNAV_TARGET_ENTITY || kind == PFNAV_TARGET_ENEMIES, "target_kind");
    PF_ARG(nents == 0 || ents, "ents");
    PF_ARG(kind != PFNAV_TARGET_ENTITY || nents == 1, "TARGET_ENTITY takes exactly one entity");
    PF_ARG(nents < (1u << 24), "nents");
    return PFNAV_OK;
}

extern "C" int pfnav_entity_seeds(pfnav_ctx *ctx, int ref_layer, int target_kind, const pfnav_footprint *ents, size_t nents,
                                  int chunk_r, int chunk_c, int32_t *out_rc, size_t cap, size_t *out_n)
{
    PF_ARG(ctx && !ctx->h_cost.empty(), "map not created");
    PF_ARG(chunk_r >= 0 && chunk_r < ctx->chunk_h && chunk_c >= 0 && chunk_c < ctx->chunk_w && out_n, "chunk / out_n");
    int rc = entity_args_ok(ctx, ref_layer, target_kind, ents, nents);
    if (rc) return rc;
    int dim = 0;
    rc = zone_dim(ctx, &dim);
    if (rc) return rc;
    std::vector<int32_t> seeds;
    rc = entity_seeds(ctx, ref_layer, target_kind, ents, nents, dim, chunk_r, chunk_c, seeds);
    if (rc) return rc;
    *out_n = seeds.size() / 2;
    if (out_rc) memcpy(out_rc, seeds.data(), std::min(cap * 2, seeds.size()) * sizeof(int32_t));
    return PFNAV_OK;
}

extern "C" int pfnav_entity_fields(pfnav_ctx *ctx, int layer, int ref_layer, int target_kind, const pfnav_footprint *ents,
                                   size_t nents, const int32_t *chunks_rc, size_t n, uint8_t *out_fields)
{
    PF_ARG(ctx && ctx->d_cost, "map not created");
    PF_NEED_DEVICE(ctx);
    if (n == 0) return PFNAV_OK;
    PF_ARG(chunks_rc && out_fields, "null buffer");
    PF_ARG(layer >= 0 && layer < ctx->nlayers, "layer");
    PF_ARG(n < (1u << 24), "n");
    int rc = entity_args_ok(ctx, ref_layer, target_kind, ents, nents);
    if (rc) return rc;
    int dim = 0;
    rc = zone_dim(ctx, &dim);
    if (rc) return rc;
    std::vector<int32_t> seeds;
    std::vector<size_t> off(n + 1, 0);
    for (size_t i = 0; i < n; i++) {
        const int cr = chunks_rc[2 * i], cc = chunks_rc[2 * i + 1];
        PF_ARG(cr >= 0 && cr < ctx->chunk_h && cc >= 0 && cc < ctx->chunk_w, "entity field: chunk");
        rc = entity_seeds(ctx, ref_layer, target_kind, ents, nents, dim, cr, cc, seeds);
        if (rc) return rc;
        off[i + 1] = seeds.size() / 2;
    }
    PF_CUDA(cudaSetDevice(ctx->device));
    uint8_t *d_out = nullptr;
    PF_CUDA(cudaMalloc(&d_out, n * 4096));
    rc = chunk_fields_launch(ctx, layer, dim, chunks_rc, n, seeds, off, d_out, ctx->tick_stream);
    if (rc == 0 && cudaMemcpy(out_fields, d_out, n * 4096, cudaMemcpyDeviceToHost) != cudaSuccess) {
        pfnav_set_error("pfnav_entity_fields: copy back failed"); rc = PFNAV_ERR_CUDA;
    }
    cudaFree(d_out);
    return rc;
}


// ------------------------------------------------------------------------------------------
// TARGET_ENEMIES / TARGET_ENTITY fields as pool destinations: what N_RequestAsyncEnemySeekField /
// N_RequestAsyncSurroundField + N_AwaitAsyncFields leave in the field cache for the chunks the seekers stand on
// (nav.c:3769-3960), consumed by N_DesiredEnemySeekVelocity / N_DesiredSurroundVelocity (k_desired_velocity).
// ------------------------------------------------------------------------------------------
extern "C" int pfnav_pool_request_entity_fields(pfnav_ctx *ctx, int dest, int layer, int ref_layer, int target_kind,
                                                const pfnav_footprint *ents, size_t nents, const int32_t *chunks_rc, size_t n,
                                                void *stream)
{
    PF_ARG(ctx && ctx->d_pool_slot, "pool not created");
    PF_NEED_DEVICE(ctx);
    PF_ARG(dest >= 0 && dest < ctx->pool_ndests, "dest");
    PF_ARG(layer >= 0 && layer < ctx->nlayers, "layer");
    int rc = entity_args_ok(ctx, ref_layer, target_kind, ents, nents);
    if (rc) return rc;
    // the footprints stay with the destination: the in-place repairs of its fields start from them (field.c:2334-2360)
    if (ctx->aux.size() < (size_t)ctx->pool_ndests) ctx->aux.resize(ctx->pool_ndests);
    pfnav_ctx::aux_target &A = ctx->aux[dest];
    A.kind = target_kind; A.layer = layer; A.ref_layer = ref_layer; A.ents.assign(ents, ents + nents);
    if (n == 0) return PFNAV_OK;
    PF_ARG(chunks_rc, "chunks");
    int dim = 0;
    if ((rc = zone_dim(ctx, &dim))) return rc;
    const int chunks = ctx->chunk_w * ctx->chunk_h;
    std::vector<int32_t> seeds;
    std::vector<size_t> off(n + 1, 0), keys(n);
    for (size_t i = 0; i < n; i++) {
        const int cr = chunks_rc[2 * i], cc = chunks_rc[2 * i + 1];
        PF_ARG(cr >= 0 && cr < ctx->chunk_h && cc >= 0 && cc < ctx->chunk_w, "entity field: chunk");
        if ((rc = entity_seeds(ctx, ref_layer, target_kind, ents, nents, dim, cr, cc, seeds))) return rc;
        off[i + 1] = seeds.size() / 2;
        keys[i] = (size_t)dest * chunks + cr * ctx->chunk_w + cc;
    }
    PF_CUDA(cudaSetDevice(ctx->device));
    PF_CUDA(pf_fields_sync(ctx));
    std::vector<int32_t> slots(n);
    bool evicted = false;
    if ((rc = pf_pool_reserve(ctx, keys.data(), n, slots.data(), &evicted))) return rc;
    cudaStream_t st = pf_stream(ctx, stream);
    uint8_t *d_out = nullptr;
    PF_CUDA(cudaMalloc(&d_out, n * 4096));
    rc = chunk_fields_launch(ctx, layer, dim, chunks_rc, n, seeds, off, d_out, st);
    for (size_t i = 0; i < n && rc == 0; i++) {
        if (cudaMemcpyAsync(ctx->d_pool_flow + (size_t)slots[i] * 4096, d_out + i * 4096, 4096, cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
            pfnav_set_error("pfnav_pool_request_entity_fields: copy into the pool failed"); rc = PFNAV_ERR_CUDA;
        }
        ctx->h_pool_has[slots[i]] |= 1;
        pfnav_field_req q;
        memset(&q, 0, sizeof(q));
        q.chunk_r = chunks_rc[2 * i]; q.chunk_c = chunks_rc[2 * i + 1]; q.layer = layer; q.faction_id = PFNAV_FACTION_ID_NONE;
        q.target_type = 2 + target_kind;          // beyond the chunk-local kinds: the frontier comes from ctx->aux[dest]
        q._pad = dest;
        ctx->h_pool_req[slots[i]] = q;
    }
    if (rc == 0) {
        cudaError_t e = cudaMemcpyAsync(ctx->d_pool_slot, ctx->h_pool_slot.data(), ctx->h_pool_slot.size() * 4, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->d_pool_los + (size_t)ctx->pool_max * 4096, ctx->h_pool_has.data(), ctx->pool_max, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { pfnav_set_error("pfnav_pool_request_entity_fields: %s", cudaGetErrorString(e)); rc = PFNAV_ERR_CUDA; }
    }
    cudaFree(d_out);
    ctx->goal_batch.valid = false;
    return rc;
}

// the chunk-local frontier of a TARGET_ENEMIES / TARGET_ENTITY pool destination (field_enemies_initial_frontier /
// field_entity_initial_frontier with the chunk itself as the region, field.c:2334-2360): tile indices r * 64 + c
int pfnav_aux_chunk_seeds(pfnav_ctx *ctx, int dest, int chunk_r, int chunk_c, std::vector<int> &out)
{
    PF_ARG(dest >= 0 && (size_t)dest < ctx->aux.size() && ctx->aux[dest].kind >= 0, "no entity fields were requested for this destination");
    const pfnav_ctx::aux_target &A = ctx->aux[dest];
    std::vector<int32_t> seeds;
    int rc = entity_seeds(ctx, A.ref_layer, A.kind, A.ents.data(), A.ents.size(), 64, chunk_r, chunk_c, seeds);
    if (rc) return rc;
    for (size_t i = 0; i + 1 < seeds.size(); i += 2) out.push_back((seeds[i] - chunk_r * 64) * 64 + (seeds[i + 1] - chunk_c * 64));
    return PFNAV_OK;
}
