// pfnav_blockers.cu -- dynamic obstacles: host-side restatement of the blocker reference counting and
// of N_Update, pushing only the chunks that changed to the device (deltas, not whole-state re-uploads
// as the reference's GLSL path does every tick, src/game/movement.c:3894-3900).
//
// Restates (reference file:line):
//   N_BlockersIncref / N_BlockersDecref                src/navigation/nav.c:4663-4683
//   n_update_blockers_circle_{ground,water,air}        src/navigation/nav.c:1051-1127
//   n_update_blockers                                  src/navigation/nav.c:1017-1049
//   M_Tile_AllUnderCircle / M_Tile_Contour             src/map/tile.c:687-718, 759-852
//   C_CircleRectIntersection, C_PointInsideRect2D, C_LineCircleIntersection
//                                                      src/phys/collision.c:997-1027, 756-768, 960-996
//   N_Update + N_ApplyDeferredInvalidations            src/navigation/nav.c:2119-2223
//   N_FC_InvalidateAllAtChunk / ...ThroughChunk        src/navigation/fieldcache.c:460-472, 481-545
#include "pfnav_internal.cuh"
#include <algorithm>
#include <limits.h>
#include <math.h>
#include <set>
#include <string.h>
#include <unordered_map>

int pfnav_route_refresh_edges(pfnav_ctx *ctx, int layer, int chunk);      // pfnav_route.cu; -1 if routing not built

namespace {

struct td { int chunk_r, chunk_c, tile_r, tile_c; };
struct v2f { float x, z; };
#define EPS_COLL (1.0f / 1024.0f)          // collision.c:64

static bool point_inside_rect(v2f p, v2f a, v2f b, v2f d)
{
    const v2f ap = {p.x - a.x, p.z - a.z}, ab = {b.x - a.x, b.z - a.z}, ad = {d.x - a.x, d.z - a.z};
    const float ap_ab = ap.x * ab.x + ap.z * ab.z, ap_ad = ap.x * ad.x + ap.z * ad.z;
    return (ap_ab >= 0.0f && ap_ab <= ab.x * ab.x + ab.z * ab.z) && (ap_ad >= 0.0f && ap_ad <= ad.x * ad.x + ad.z * ad.z);
}

static bool line_circle(float ax, float az, float bx, float bz, v2f c, float radius)
{
    const float dx = bx - ax, dz = bz - az;
    const float A = pow(dx, 2) + pow(dz, 2);
    const float B = 2 * (dx * (ax - c.x) + dz * (az - c.z));
    const float C = pow(ax - c.x, 2) + pow(az - c.z, 2) - pow(radius, 2);
    const float det = pow(B, 2) - (4 * A * C);
    float t;
    if (det < 0.0f || A < EPS_COLL) return false;
    else if (det == 0.0f) t = -B / (2 * A);
    else {
        const float t1 = (-B + sqrt(det)) / (2 * A), t2 = (-B - sqrt(det)) / (2 * A);
        t = std::min(t1, t2);
    }
    if (t < 0.0f || t > 1.0f) return false;
    return true;
}

// C_CircleRectIntersection (collision.c:997); rect = {x, z, width, height}, x decreasing with the column
static bool circle_rect(v2f center, float radius, float rx, float rz, float w, float h)
{
    const v2f corners[4] = {{rx - w, rz}, {rx, rz}, {rx, rz + h}, {rx - w, rz + h}};
    if (point_inside_rect(center, corners[0], corners[1], corners[3])) return true;
    for (int i = 0; i < 4; i++) {
        const float ddx = corners[i].x - center.x, ddz = corners[i].z - center.z;
        if ((float)sqrt(ddx * ddx + ddz * ddz) <= radius) return true;
    }
    for (int i = 0; i < 4; i++) {
        const v2f a = corners[i], b = corners[(i + 1) & 3];
        if (line_circle(a.x, a.z, b.x, b.z, center, radius)) return true;
    }
    return false;
}

// M_Tile_DescForPoint2D with the nav resolution (tile.c:547)
static bool desc_for_point(const pfnav_ctx *ctx, float px, float pz, td *out)
{
    const float width = (float)(ctx->chunk_w * 256), height = (float)(ctx->chunk_h * 256);
    if (px > ctx->map_x || px < ctx->map_x - width) return false;
    if (pz < ctx->map_z || pz > ctx->map_z + height) return false;
    int chunk_r = (int)(fabs(ctx->map_z - pz) / 256.0f), chunk_c = (int)(fabs(ctx->map_x - px) / 256.0f);
    chunk_r = std::min(std::max(chunk_r, 0), ctx->chunk_h - 1);
    chunk_c = std::min(std::max(chunk_c, 0), ctx->chunk_w - 1);
    const float bx = ctx->map_x - (chunk_c * 256.0f), bz = ctx->map_z + (chunk_r * 256.0f);
    int tile_r = (int)(fabs(bz - pz) / 4), tile_c = (int)(fabs(bx - px) / 4);
    out->chunk_r = chunk_r; out->chunk_c = chunk_c;
    out->tile_r = std::min(std::max(tile_r, 0), 63); out->tile_c = std::min(std::max(tile_c, 0), 63);
    return true;
}

// M_Tile_AllUnderCircle (tile.c:687)
static size_t tiles_under_circle(const pfnav_ctx *ctx, v2f c, float radius, td *out, size_t maxout)
{
    td tile;
    if (!desc_for_point(ctx, c.x, c.z, &tile)) return 0;
    const int ntiles = (int)ceil(radius / 4);
    size_t ret = 0;
    for (int dr = -ntiles; dr <= ntiles; dr++)
        for (int dc = -ntiles; dc <= ntiles; dc++) {
            const int ar = tile.chunk_r * 64 + tile.tile_r + dr, ac = tile.chunk_c * 64 + tile.tile_c + dc;
            if (ar < 0 || ar >= ctx->chunk_h * 64 || ac < 0 || ac >= ctx->chunk_w * 64) continue;
            const td cur = {ar / 64, ac / 64, ar % 64, ac % 64};
            // M_Tile_Bounds (tile.c:356)
            const float bx = (ctx->map_x - (float)(cur.chunk_c * 256)) - (float)(cur.tile_c * 4);
            const float bz = (ctx->map_z + (float)(cur.chunk_r * 256)) + (float)(cur.tile_r * 4);
            if (!circle_rect(c, radius, bx, bz, 4.0f, 4.0f)) continue;
            out[ret++] = cur;
            if (ret == maxout) return ret;
        }
    return ret;
}

// M_Tile_LineSupercoverTilesSorted (tile.c:430), for a segment that STARTS inside the map (the only
// case pfnav_blockers_*_obb accepts): Amanatides-Woo traversal in the reference's float arithmetic.
static size_t line_supercover(const pfnav_ctx *ctx, float ax, float az, float bx, float bz, td *out, size_t maxout)
{
    size_t ret = 0;
    if (maxout == 0) return 0;
    float dx = bx - ax, dz = bz - az;
    const float len = (float)sqrt(dx * dx + dz * dz);
    dx = dx / len; dz = dz / len;                                  // PFM_Vec2_Normal
    td cur;
    if (!desc_for_point(ctx, ax, az, &cur)) return 0;
    const int step_c = dx <= 0.0f ? 1 : -1;
    const int step_r = dz >= 0.0f ? 1 : -1;
    const float t_delta_x = (float)fabs(4 / dx), t_delta_z = (float)fabs(4 / dz);       // TILE_X_DIM is an int
    const float bnx = (ctx->map_x - (float)(cur.chunk_c * 256)) - (float)(cur.tile_c * 4);
    const float bnz = (ctx->map_z + (float)(cur.chunk_r * 256)) + (float)(cur.tile_r * 4);
    float t_max_x = (step_c > 0) ? (float)(fabs(ax - (bnx - 4.0f)) / fabs(dx)) : (float)(fabs(ax - bnx) / fabs(dx));
    float t_max_z = (step_r > 0) ? (float)(fabs(az - (bnz + 4.0f)) / fabs(dz)) : (float)(fabs(az - bnz) / fabs(dz));
    td fin;
    const bool ends_inside = desc_for_point(ctx, bx, bz, &fin);
    do {
        out[ret++] = cur;
        int dc = 0, dr = 0;
        if (t_max_x < t_max_z) { t_max_x = t_max_x + t_delta_x; dc = step_c; }
        else                   { t_max_z = t_max_z + t_delta_z; dr = step_r; }
        if (ends_inside && cur.chunk_r == fin.chunk_r && cur.chunk_c == fin.chunk_c && cur.tile_r == fin.tile_r && cur.tile_c == fin.tile_c)
            break;
        const int ar = cur.chunk_r * 64 + cur.tile_r + dr, ac = cur.chunk_c * 64 + cur.tile_c + dc;      // M_Tile_RelativeDesc
        if (ar < 0 || ar >= ctx->chunk_h * 64 || ac < 0 || ac >= ctx->chunk_w * 64) break;
        cur = {ar / 64, ac / 64, ar % 64, ac % 64};
    } while (ret < maxout);
    return ret;
}

// M_Tile_AllUnderObj (tile.c:594): supercover of the four bottom edges (duplicates at the corners included, as
// in the reference: the refcounts see them twice) plus every tile of the bounding tile box -- upper bounds
// EXCLUDED, tile.c:656-659 -- whose centre lies inside the rectangle (C_PointInsideRect2D, collision.c:756).
// c[4] = bottom corners obb->corners[0], [1], [5], [4] as (x, z).
static size_t tiles_under_obb(const pfnav_ctx *ctx, const v2f c[4], td *out, size_t maxout)
{
    size_t ret = 0;
    int min_r = ctx->chunk_h * 64 - 1, max_r = 0, min_c = ctx->chunk_w * 64 - 1, max_c = 0;
    // the reference starts its column minimum at {chunk_w-1, chunk_w-1} (tile.c:624): reproduce it
    int min_c_init = (ctx->chunk_w - 1) * 64 + (ctx->chunk_w - 1);
    min_c = min_c_init;
    for (int i = 0; i < 4; i++) {
        const v2f a = c[i], b = c[(i + 1) & 3];
        const size_t n = line_supercover(ctx, a.x, a.z, b.x, b.z, out + ret, maxout - ret);
        const td *d = out + ret;
        ret += n;
        if (ret == maxout) return ret;
        for (size_t j = 0; j < n; j++) {
            const int ar = d[j].chunk_r * 64 + d[j].tile_r, ac = d[j].chunk_c * 64 + d[j].tile_c;
            min_r = std::min(min_r, ar); max_r = std::max(max_r, ar);
            min_c = std::min(min_c, ac); max_c = std::max(max_c, ac);
        }
    }
    const v2f ab = {c[1].x - c[0].x, c[1].z - c[0].z}, ad = {c[3].x - c[0].x, c[3].z - c[0].z};
    const float abab = ab.x * ab.x + ab.z * ab.z, adad = ad.x * ad.x + ad.z * ad.z;
    for (int r = min_r; r < max_r; r++)
        for (int cc = min_c; cc < max_c; cc++) {
            const float bx = (ctx->map_x - (float)((cc / 64) * 256)) - (float)((cc % 64) * 4);
            const float bz = (ctx->map_z + (float)((r / 64) * 256)) + (float)((r % 64) * 4);
            const v2f ctr = {bx - 4.0f / 2.0f, bz + 4.0f / 2.0f};
            const v2f ap = {ctr.x - c[0].x, ctr.z - c[0].z};
            const float apab = ap.x * ab.x + ap.z * ab.z, apad = ap.x * ad.x + ap.z * ad.z;
            if ((apab >= 0.0f && apab <= abab) && (apad >= 0.0f && apad <= adad)) {
                out[ret++] = {r / 64, cc / 64, r % 64, cc % 64};
                if (ret == maxout) return ret;
            }
        }
    return ret;
}

// M_Tile_Contour (tile.c:759)
static size_t tiles_contour(const pfnav_ctx *ctx, size_t ntds, const td *tds, td *out, size_t maxout)
{
    if (ntds == 0) return 0;
    int minr = INT_MAX, minc = INT_MAX, maxr = INT_MIN, maxc = INT_MIN;
    for (size_t i = 0; i < ntds; i++) {
        const int ar = tds[i].chunk_r * 64 + tds[i].tile_r, ac = tds[i].chunk_c * 64 + tds[i].tile_c;
        minr = std::min(minr, ar); minc = std::min(minc, ac); maxr = std::max(maxr, ar); maxc = std::max(maxc, ac);
    }
    const int dr = maxr - minr + 1, dc = maxc - minc + 1;
    const size_t width = dc + 2, height = dr + 2;
    std::vector<uint8_t> marked(width * height, 0);
    for (size_t i = 0; i < ntds; i++) {
        const int ar = tds[i].chunk_r * 64 + tds[i].tile_r, ac = tds[i].chunk_c * 64 + tds[i].tile_c;
        marked[(ar - minr + 1) * width + (ac - minc + 1)] = 1;
    }
    size_t ret = 0;
    for (int r = minr - 1; r <= maxr + 1; r++)
        for (int c = minc - 1; c <= maxc + 1; c++) {
            if (r < 0 || r >= ctx->chunk_h * 64 || c < 0 || c >= ctx->chunk_w * 64) continue;
            const int relr = r - minr + 1, relc = c - minc + 1;
            if (marked[relr * width + relc]) continue;
            if (ret == maxout) return ret;
            bool contour = false;
            if ((relr > 0 && marked[(relr - 1) * (dc + 2) + relc]) || (relr < dr && marked[(relr + 1) * (dc + 2) + relc]) ||
                (relc > 0 && marked[relr * (dc + 2) + (relc - 1)]) || (relc < dc && marked[relr * (dc + 2) + (relc + 1)]))
                contour = true;
            if ((relr > 0 && relc > 0 && marked[(relr - 1) * (dc + 2) + (relc - 1)]) ||
                (relr > 0 && relc < dc && marked[(relr - 1) * (dc + 2) + (relc + 1)]) ||
                (relr < dr && relc > 0 && marked[(relr + 1) * (dc + 2) + (relc - 1)]) ||
                (relr < dr && relc < dc && marked[(relr + 1) * (dc + 2) + (relc + 1)]))
                contour = true;
            if (contour) out[ret++] = {r / 64, c / 64, r % 64, c % 64};
        }
    return ret;
}

// dirty (layer, chunk) sets live in the context: pfnav_ctx::dirty (occupancy changed), ::fdirty (faction mask changed)

// n_update_blockers (nav.c:1017) on the host mirror; layers the context does not hold are skipped
static void apply(pfnav_ctx *ctx, int layer, int faction_id, const td *tds, size_t n, int delta)
{
    if (layer >= ctx->nlayers) return;
    const size_t chunks = (size_t)ctx->chunk_w * ctx->chunk_h, ltiles = chunks * 4096;
    const size_t lbase = (size_t)layer * ltiles;
    const bool fac = faction_id >= 0 && faction_id < 15;
    if (fac) {      // chunk->factions[faction_id] (nav.c:1032), allocated on first use
        if (ctx->h_fac.size() < (size_t)ctx->nlayers) ctx->h_fac.resize(ctx->nlayers);
        if (ctx->h_fac[layer].empty()) ctx->h_fac[layer].assign(ltiles * 15, 0);
        if (ctx->h_fmask.size() < ltiles * ctx->nlayers) ctx->h_fmask.assign(ltiles * ctx->nlayers, 0);
    }
    for (size_t i = 0; i < n; i++) {
        const int chunk = tds[i].chunk_r * ctx->chunk_w + tds[i].chunk_c;
        const int t = tds[i].tile_r * 64 + tds[i].tile_c;
        uint16_t &v = ctx->h_blk[lbase + (size_t)chunk * 4096 + t];
        const int prev = v;
        v = (uint16_t)(prev + delta);
        if (!!v != !!prev) ctx->dirty.insert({layer, chunk});
        if (fac) {
            uint8_t &fv = ctx->h_fac[layer][((size_t)chunk * 15 + faction_id) * 4096 + t];
            const int fprev = fv;
            fv = (uint8_t)(fprev + delta);
            if (!!fv != !!fprev) {
                uint16_t &m = ctx->h_fmask[lbase + (size_t)chunk * 4096 + t];
                m = fv ? (uint16_t)(m | (1u << faction_id)) : (uint16_t)(m & ~(1u << faction_id));
                ctx->fdirty.insert({layer, chunk});
            }
        }
    }
}

static int blockers_circle(pfnav_ctx *ctx, float x, float z, float range, int faction_id, uint32_t flags, int delta)
{
    td tds[1024], o3[1024], o5[1024], o7[1024];
    const size_t n = tiles_under_circle(ctx, {x, z}, range, tds, 1024);
    const size_t n3 = tiles_contour(ctx, n, tds, o3, 1024);
    const size_t n5 = tiles_contour(ctx, n3, o3, o5, 1024);
    const size_t n7 = tiles_contour(ctx, n5, o5, o7, 1024);
    // layer groups: ground 0..3, water 4..7, air 8..11 (nav.h:78-92); non-air entities block ground AND water
    const int groups[2] = {(flags & PFNAV_FLAG_AIR) ? 8 : 4, (flags & PFNAV_FLAG_AIR) ? -1 : 0};
    for (int gi = 0; gi < 2; gi++) {
        const int g = groups[gi];
        if (g < 0) continue;
        apply(ctx, g + 0, faction_id, tds, n, delta);
        apply(ctx, g + 1, faction_id, tds, n, delta); apply(ctx, g + 1, faction_id, o3, n3, delta);
        apply(ctx, g + 2, faction_id, tds, n, delta); apply(ctx, g + 2, faction_id, o3, n3, delta); apply(ctx, g + 2, faction_id, o5, n5, delta);
        apply(ctx, g + 3, faction_id, tds, n, delta); apply(ctx, g + 3, faction_id, o3, n3, delta); apply(ctx, g + 3, faction_id, o5, n5, delta); apply(ctx, g + 3, faction_id, o7, n7, delta);
    }
    return PFNAV_OK;
}

// n_update_blockers_obb_{ground,water,air} (nav.c:1135-1211)
static int blockers_obb(pfnav_ctx *ctx, const float *corners_xz, int faction_id, uint32_t flags, int delta)
{
    v2f c[4];
    for (int i = 0; i < 4; i++) {
        c[i] = {corners_xz[2 * i], corners_xz[2 * i + 1]};
        td t;
        if (!desc_for_point(ctx, c[i].x, c[i].z, &t)) {
            pfnav_set_error("pfnav_blockers_*_obb: corner %d lies outside the map (clipped boxes are not supported)", i);
            return PFNAV_ERR_ARG;
        }
    }
    td tds[1024], o3[1024], o5[1024], o7[1024];
    const size_t n = tiles_under_obb(ctx, c, tds, 1024);
    const size_t n3 = tiles_contour(ctx, n, tds, o3, 1024);
    const size_t n5 = tiles_contour(ctx, n3, o3, o5, 1024);
    const size_t n7 = tiles_contour(ctx, n5, o5, o7, 1024);
    const int groups[2] = {(flags & PFNAV_FLAG_AIR) ? 8 : 4, (flags & PFNAV_FLAG_AIR) ? -1 : 0};
    for (int gi = 0; gi < 2; gi++) {
        const int g = groups[gi];
        if (g < 0) continue;
        apply(ctx, g + 0, faction_id, tds, n, delta);
        apply(ctx, g + 1, faction_id, tds, n, delta); apply(ctx, g + 1, faction_id, o3, n3, delta);
        apply(ctx, g + 2, faction_id, tds, n, delta); apply(ctx, g + 2, faction_id, o3, n3, delta); apply(ctx, g + 2, faction_id, o5, n5, delta);
        apply(ctx, g + 3, faction_id, tds, n, delta); apply(ctx, g + 3, faction_id, o3, n3, delta); apply(ctx, g + 3, faction_id, o5, n5, delta); apply(ctx, g + 3, faction_id, o7, n7, delta);
    }
    return PFNAV_OK;
}

}   // namespace

// The tiles an entity occupies as a field target (field_entity_initial_frontier field.c:1334-1355,
// field_enemies_initial_frontier :1262-1283): circle (selection radius) or building OBB footprint, then `rings`
// contour rings, each taken around everything collected so far, all inside one 512-entry array.
// out_rc: up to 512 (r, c) pairs. Returns the count, or -1 when an OBB corner lies outside the map.
int pfnav_footprint_tiles(const pfnav_ctx *ctx, const pfnav_footprint *e, int rings, int32_t *out_rc)
{
    td tds[512];
    size_t n;
    if (e->is_building) {
        v2f c[4];
        for (int i = 0; i < 4; i++) {
            c[i] = {e->corners_xz[2 * i], e->corners_xz[2 * i + 1]};
            td t;
            if (!desc_for_point(ctx, c[i].x, c[i].z, &t)) return -1;
        }
        n = tiles_under_obb(ctx, c, tds, 512);
    } else {
        n = tiles_under_circle(ctx, {e->x, e->z}, e->sel_radius, tds, 512);
    }
    for (int k = 0; k < rings; k++) n += tiles_contour(ctx, n, tds, tds + n, 512 - n);
    for (size_t i = 0; i < n; i++) { out_rc[2 * i] = tds[i].chunk_r * 64 + tds[i].tile_r; out_rc[2 * i + 1] = tds[i].chunk_c * 64 + tds[i].tile_c; }
    return (int)n;
}

void pfnav_blockers_forget(pfnav_ctx *ctx) { ctx->dirty.clear(); ctx->fdirty.clear(); }

extern "C" int pfnav_blockers_incref(pfnav_ctx *ctx, float x, float z, float range, int faction_id, uint32_t flags)
{
    PF_ARG(ctx && ctx->d_cost, "map not created");
    return blockers_circle(ctx, x, z, range, faction_id, flags, +1);
}

// N_BlockersIncrefOBB / N_BlockersDecrefOBB (nav.c:4685-4705). corners_xz: the bottom face of the box,
// obb->corners[0], [1], [5], [4] as x,z pairs (tile.c:599).
extern "C" int pfnav_blockers_incref_obb(pfnav_ctx *ctx, const float *corners_xz, int faction_id, uint32_t flags)
{
    PF_ARG(ctx && ctx->d_cost && corners_xz, "map not created / null");
    return blockers_obb(ctx, corners_xz, faction_id, flags, +1);
}

extern "C" int pfnav_blockers_decref_obb(pfnav_ctx *ctx, const float *corners_xz, int faction_id, uint32_t flags)
{
    PF_ARG(ctx && ctx->d_cost && corners_xz, "map not created / null");
    return blockers_obb(ctx, corners_xz, faction_id, flags, -1);
}

extern "C" int pfnav_blockers_decref(pfnav_ctx *ctx, float x, float z, float range, int faction_id, uint32_t flags)
{
    PF_ARG(ctx && ctx->d_cost, "map not created");
    return blockers_circle(ctx, x, z, range, faction_id, flags, -1);
}

// A tick's worth of N_BlockersIncref / N_BlockersDecref calls (nav.c:4663-4683) in one call, applied in order.
extern "C" int pfnav_blockers_batch(pfnav_ctx *ctx, const pfnav_blocker_op *ops, size_t n)
{
    PF_ARG(ctx && ctx->d_cost && (n == 0 || ops), "map not created / null");
    for (size_t i = 0; i < n; i++) {
        PF_ARG(ops[i].delta == 1 || ops[i].delta == -1, "blocker op: delta must be +1 or -1");
        int rc = blockers_circle(ctx, ops[i].x, ops[i].z, ops[i].range, ops[i].faction_id, ops[i].flags, ops[i].delta);
        if (rc) return rc;
    }
    return PFNAV_OK;
}

// N_Update + N_ApplyDeferredInvalidations: recompute the local islands of every dirty chunk, refresh the
// portal edge states there, push the chunk (blockers + islands) to the device, and invalidate pool
// entries: everything AT a dirty chunk; and, when an edge state flipped, every field of every
// destination whose path runs THROUGH that chunk. *out_ndirty = number of (layer, chunk) pairs handled.
extern "C" int pfnav_map_commit(pfnav_ctx *ctx, int *out_ndirty)
{
    PF_ARG(ctx && ctx->d_cost, "map not created");
    if (ctx->device >= 0) {
        PF_CUDA(cudaSetDevice(ctx->device));
        PF_CUDA(pf_fields_sync(ctx));     // forked LOS chains still read the map that is about to change
        // ticks / field launches on the context stream or on caller streams (all non-blocking, so not ordered against
        // the synchronous copies below) may still read the grids
        PF_CUDA(cudaDeviceSynchronize());
    }
    {   // faction masks follow the refcounts at once (they only matter to attacking requests)
        for (const auto &lc : ctx->fdirty) { int rc = pfnav_fmask_push_chunk(ctx, lc.first, lc.second); if (rc) return rc; }
        if (!ctx->fdirty.empty()) ctx->map_epoch++;
        ctx->fdirty.clear();
    }
    int nd = 0;
    if (!ctx->dirty.empty()) {
        const int chunks = ctx->chunk_w * ctx->chunk_h;
        bool pool_touched = false;
        for (const auto &lc : ctx->dirty) {
            const int layer = lc.first, chunk = lc.second;
            int rc = pfnav_map_refresh_chunk(ctx, layer, chunk / ctx->chunk_w, chunk % ctx->chunk_w);
            if (rc) return rc;
            const int flipped = pfnav_route_refresh_edges(ctx, layer, chunk);
            nd++;
            if (!ctx->h_pool_slot.empty()) {
                for (int d = 0; d < ctx->pool_ndests; d++) {
                    const size_t si = (size_t)d * chunks + chunk;
                    const int slot = ctx->h_pool_slot[si];
                    if (slot < 0 || !ctx->h_pool_has[slot]) continue;
                    pool_touched = true;
                    if (flipped > 0) {
                        // N_FC_InvalidateAllThroughChunk: the whole path of this destination goes
                        for (int c2 = 0; c2 < chunks; c2++) {
                            const int s2 = ctx->h_pool_slot[(size_t)d * chunks + c2];
                            if (s2 >= 0) ctx->h_pool_has[s2] = 0;
                            ctx->h_pool_ffid[(size_t)d * chunks + c2] = 0;
                        }
                    } else {
                        ctx->h_pool_has[slot] = 0;
                        ctx->h_pool_ffid[si] = 0;
                    }
                }
            }
        }
        ctx->dirty.clear();
        if (pool_touched && ctx->device >= 0) {
            PF_CUDA(cudaSetDevice(ctx->device));
            PF_CUDA(cudaMemcpy(ctx->d_pool_los + (size_t)ctx->pool_max * 4096, ctx->h_pool_has.data(), ctx->pool_max,
                               cudaMemcpyHostToDevice));
        }
        ctx->goal_batch.valid = false;
    }
    if (out_ndirty) *out_ndirty = nd;
    return PFNAV_OK;
}

// Read back one layer's blocker counts ([chunk][64][64] u16) from the host mirror (parity tests).
extern "C" int pfnav_blockers_get(pfnav_ctx *ctx, int layer, uint16_t *out)
{
    PF_ARG(ctx && out && layer >= 0 && layer < ctx->nlayers, "args");
    const size_t ltiles = (size_t)ctx->chunk_w * ctx->chunk_h * 4096;
    memcpy(out, ctx->h_blk.data() + ltiles * layer, ltiles * 2);
    return PFNAV_OK;
}

// Read back one layer's per-faction blocker counts (nav_chunk::factions, nav_data.h: [chunk][15][64][64] u8)
// from the host mirror; all zero until a faction-tagged blocker was counted or pfnav_map_upload_factions ran.
extern "C" int pfnav_blockers_get_factions(pfnav_ctx *ctx, int layer, uint8_t *out)
{
    PF_ARG(ctx && out && layer >= 0 && layer < ctx->nlayers, "args");
    const size_t n = (size_t)ctx->chunk_w * ctx->chunk_h * 15 * 4096;
    if ((size_t)layer < ctx->h_fac.size() && ctx->h_fac[layer].size() == n) memcpy(out, ctx->h_fac[layer].data(), n);
    else memset(out, 0, n);
    return PFNAV_OK;
}
