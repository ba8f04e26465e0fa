// pfnav_plan.cu -- host-side navigation structure: local islands, portals, and a goal planner
// that emits the flow/LOS field requests the device kernels consume.
//
// Restates (reference file:line):
//   n_update_local_islands / n_visit_island_local     src/navigation/nav.c:967-984, 903-950
//   n_create_portals / n_link_chunks                  src/navigation/nav.c:563-591, 477-561
// and provides a breadth-first goal planner over (chunk, local island) nodes that yields, for
// EVERY chunk connected to the goal, the same kind of TARGET_TILE / TARGET_PORTAL requests
// n_request_path (nav.c:1774-2047) issues along one portal path. It minimises hop count, not the
// reference's portal travel cost, so which portal a chunk steers to may differ from
// AStar_PortalGraphPath's choice (a_star.c:429); each emitted field is still the reference's
// field for that request. The cost-faithful search is listed under "next" in DESIGN.md.
#include "pfnav_internal.cuh"
#include <algorithm>
#include <atomic>
#include <deque>
#include <string.h>
#include <thread>

static inline size_t chunk_off(const pfnav_ctx *ctx, int layer, int cr, int cc)
{
    return ((size_t)layer * ctx->chunk_w * ctx->chunk_h + (size_t)cr * ctx->chunk_w + cc) * 4096;
}

// n_update_local_islands (nav.c:967): ids from 1 in row-major discovery order, 4-connected flood
// inside the chunk over tiles that are passable and not blocked; 0xFFFF elsewhere.
static void local_islands_chunk(const uint8_t *cost, const uint16_t *blk, uint16_t *out)
{
    for (int i = 0; i < 4096; i++) out[i] = 0xFFFF;
    uint16_t next = 0;
    std::vector<int> q;
    q.reserve(4096);
    for (int t = 0; t < 4096; t++) {
        if (out[t] != 0xFFFF || cost[t] == 0xFF || blk[t] > 0) continue;
        const uint16_t id = ++next;
        q.clear();
        q.push_back(t);
        out[t] = id;
        for (size_t h = 0; h < q.size(); h++) {
            const int cur = q[h], r = cur >> 6, c = cur & 63;
            const int nb[4] = {c > 0 ? cur - 1 : -1, c < 63 ? cur + 1 : -1, r > 0 ? cur - 64 : -1, r < 63 ? cur + 64 : -1};
            for (int e = 0; e < 4; e++) {
                const int n = nb[e];
                if (n < 0 || out[n] != 0xFFFF || cost[n] == 0xFF || blk[n] > 0) continue;
                out[n] = id;
                q.push_back(n);
            }
        }
    }
}

// n_link_chunks (nav.c:477): a portal is a maximal run of border tiles passable on both sides
// (cost only, blockers ignored); a run that would start on the last tile of the line is never
// closed and therefore dropped, exactly as the reference's loop does.
static void link_chunks(pfnav_ctx *ctx, int layer, int ar, int ac, int br, int bc, bool vertical_pair)
{
    // vertical_pair: a is above b (a's bottom row vs b's top row); else a is left of b
    auto &pa = ctx->portals[layer][ar * ctx->chunk_w + ac];
    auto &pb = ctx->portals[layer][br * ctx->chunk_w + bc];
    const uint8_t *ca = ctx->h_cost.data() + chunk_off(ctx, layer, ar, ac);
    const uint8_t *cb = ctx->h_cost.data() + chunk_off(ctx, layer, br, bc);
    bool in_portal = false;
    int start = 0;
    for (int i = 0; i < 64; i++) {
        const uint8_t va = vertical_pair ? ca[63 * 64 + i] : ca[i * 64 + 63];
        const uint8_t vb = vertical_pair ? cb[i] : cb[i * 64];
        const bool can_cross = va != 0xFF && vb != 0xFF;
        if (can_cross && !in_portal) {
            in_portal = true;
            start = i;
        } else if (in_portal && (!can_cross || i == 63)) {
            const int end = !can_cross ? i - 1 : i;
            in_portal = false;
            pfnav_ctx::portal_t A, B;
            A.chunk_r = (int16_t)ar; A.chunk_c = (int16_t)ac; B.chunk_r = (int16_t)br; B.chunk_c = (int16_t)bc;
            if (vertical_pair) {
                A.r0 = 63; A.c0 = (int16_t)start; A.r1 = 63; A.c1 = (int16_t)end;
                B.r0 = 0;  B.c0 = (int16_t)start; B.r1 = 0;  B.c1 = (int16_t)end;
            } else {
                A.r0 = (int16_t)start; A.c0 = 63; A.r1 = (int16_t)end; A.c1 = 63;
                B.r0 = (int16_t)start; B.c0 = 0;  B.r1 = (int16_t)end; B.c1 = 0;
            }
            A.conn_chunk = br * ctx->chunk_w + bc; A.conn_idx = (int32_t)pb.size();
            B.conn_chunk = ar * ctx->chunk_w + ac; B.conn_idx = (int32_t)pa.size();
            pa.push_back(A);
            pb.push_back(B);
        }
    }
}

// N_NewCtxForMapData's structural part for one layer (nav.c:2284): local islands (uploaded to the
// device for TARGET_PORTAL seeding) and the portal table. Call after pfnav_map_upload_layer.
extern "C" int pfnav_map_build_nav(pfnav_ctx *ctx, int layer)
{
    PF_ARG(ctx && ctx->d_cost, "map not created");
    PF_ARG(layer >= 0 && layer < ctx->nlayers, "layer");
    const int chunks = ctx->chunk_w * ctx->chunk_h;
    const size_t ltiles = (size_t)chunks * 4096;
    for (int ch = 0; ch < chunks; ch++)
        local_islands_chunk(ctx->h_cost.data() + ltiles * layer + (size_t)ch * 4096,
                            ctx->h_blk.data() + ltiles * layer + (size_t)ch * 4096,
                            ctx->h_liid.data() + ltiles * layer + (size_t)ch * 4096);
    // n_create_portals (nav.c:563): chunks row-major, bottom link before right link
    ctx->portals[layer].assign(chunks, {});
    for (int r = 0; r < ctx->chunk_h; r++)
        for (int c = 0; c < ctx->chunk_w; c++) {
            if (r < ctx->chunk_h - 1) link_chunks(ctx, layer, r, c, r + 1, c, true);
            if (c < ctx->chunk_w - 1) link_chunks(ctx, layer, r, c, r, c + 1, false);
        }
    for (int ch = 0; ch < chunks; ch++)
        if (ctx->portals[layer][ch].size() > 64) {
            pfnav_set_error("pfnav_map_build_nav: chunk %d has %zu portals (MAX_PORTALS_PER_CHUNK is 64, nav_data.h:44)",
                            ch, ctx->portals[layer][ch].size());
            return PFNAV_ERR_STATE;
        }
    // push the islands to the device image (host-only contexts just keep the mirrors)
    return pfnav_map_upload_layer(ctx, layer, ctx->h_cost.data() + ltiles * layer, ctx->h_blk.data() + ltiles * layer,
                                  ctx->h_liid.data() + ltiles * layer);
}

// Recompute the local islands of one chunk after its blockers changed (n_update_dirty_local_islands,
// nav.c:986) and push the chunk to the device.
extern "C" int pfnav_map_refresh_chunk(pfnav_ctx *ctx, int layer, int chunk_r, int chunk_c)
{
    PF_ARG(ctx && ctx->d_cost, "map not created");
    PF_ARG(layer >= 0 && layer < ctx->nlayers, "layer");
    PF_ARG(chunk_r >= 0 && chunk_r < ctx->chunk_h && chunk_c >= 0 && chunk_c < ctx->chunk_w, "chunk coords");
    const size_t off = chunk_off(ctx, layer, chunk_r, chunk_c);
    local_islands_chunk(ctx->h_cost.data() + off, ctx->h_blk.data() + off, ctx->h_liid.data() + off);
    return pfnav_map_update_chunk(ctx, layer, chunk_r, chunk_c, nullptr, ctx->h_blk.data() + off, ctx->h_liid.data() + off);
}

extern "C" int pfnav_local_islands_get(pfnav_ctx *ctx, int layer, uint16_t *out)
{
    PF_ARG(ctx && out && layer >= 0 && layer < ctx->nlayers, "args");
    const size_t ltiles = (size_t)ctx->chunk_w * ctx->chunk_h * 4096;
    memcpy(out, ctx->h_liid.data() + ltiles * layer, ltiles * 2);
    return PFNAV_OK;
}

// Portal table in the same 10-int row format as the oracle harness (chunk_r, chunk_c, idx, ep0.r,
// ep0.c, ep1.r, ep1.c, conn_chunk_idx, conn_portal_idx, 0). Returns the count via *out_n.
extern "C" int pfnav_portals_get(pfnav_ctx *ctx, int layer, int32_t *out, int maxout, int *out_n)
{
    PF_ARG(ctx && out_n && layer >= 0 && layer < ctx->nlayers, "args");
    PF_ARG((size_t)layer < ctx->portals.size() && !ctx->portals[layer].empty(), "pfnav_map_build_nav not called for this layer");
    int n = 0;
    for (size_t ch = 0; ch < ctx->portals[layer].size(); ch++)
        for (size_t p = 0; p < ctx->portals[layer][ch].size(); p++) {
            const auto &P = ctx->portals[layer][ch][p];
            if (out && n < maxout) {
                int32_t *o = out + (size_t)n * 10;
                o[0] = P.chunk_r; o[1] = P.chunk_c; o[2] = (int32_t)p; o[3] = P.r0; o[4] = P.c0; o[5] = P.r1; o[6] = P.c1;
                o[7] = P.conn_chunk; o[8] = P.conn_idx; o[9] = 0;
            }
            n++;
        }
    *out_n = n;
    return PFNAV_OK;
}

// Goal planner: breadth-first over (chunk, local island) nodes from the goal tile. For every node
// emits one flow request (TARGET_TILE for the goal's own node, TARGET_PORTAL towards the parent
// node otherwise; a chunk's first node has init=1, later nodes of the same chunk merge in place
// like nav.c:1998-2008) and, for a chunk's first node, one LOS request chained to its parent chunk.
// `field_slot[i]` / `los_slot[i]` give the chunk index each request's output belongs to, and
// `flow_wave[i]` the launch wave (requests that update the same chunk are serialised by wave).
extern "C" int pfnav_plan_goal(pfnav_ctx *ctx, int layer, int tgt_chunk_r, int tgt_chunk_c, int tgt_tile_r,
                               int tgt_tile_c, pfnav_field_req *flow_out, int32_t *flow_chunk, int32_t *flow_wave,
                               int max_flow, int *n_flow, pfnav_los_req *los_out, int32_t *los_chunk, int max_los,
                               int *n_los)
{
    PF_ARG(ctx && ctx->d_cost && n_flow && n_los, "args");
    PF_ARG(layer >= 0 && layer < ctx->nlayers, "layer");
    PF_ARG((size_t)layer < ctx->portals.size() && !ctx->portals[layer].empty(), "pfnav_map_build_nav not called for this layer");
    PF_ARG(tgt_chunk_r >= 0 && tgt_chunk_r < ctx->chunk_h && tgt_chunk_c >= 0 && tgt_chunk_c < ctx->chunk_w &&
           tgt_tile_r >= 0 && tgt_tile_r < 64 && tgt_tile_c >= 0 && tgt_tile_c < 64, "target tile");
    const int cw = ctx->chunk_w, chunks = cw * ctx->chunk_h;
    const size_t ltiles = (size_t)chunks * 4096;
    const uint16_t *liid = ctx->h_liid.data() + ltiles * layer;
    *n_flow = 0; *n_los = 0;
    struct node { int chunk; uint16_t li; };
    std::deque<node> q;
    std::vector<std::vector<uint16_t>> seen(chunks);      // islands already planned per chunk
    std::vector<int> los_index(chunks, -1), nodes_in_chunk(chunks, 0);
    const int tchunk = tgt_chunk_r * cw + tgt_chunk_c;
    const uint16_t tli = liid[(size_t)tchunk * 4096 + tgt_tile_r * 64 + tgt_tile_c];
    auto emit_flow = [&](const pfnav_field_req &rq, int chunk) -> bool {
        if (*n_flow >= max_flow) return false;
        flow_out[*n_flow] = rq;
        flow_out[*n_flow].init = nodes_in_chunk[chunk] == 0;
        flow_chunk[*n_flow] = chunk;
        flow_wave[*n_flow] = nodes_in_chunk[chunk];
        nodes_in_chunk[chunk]++;
        (*n_flow)++;
        return true;
    };
    pfnav_field_req base;
    memset(&base, 0, sizeof(base));
    base.layer = layer; base.faction_id = PFNAV_FACTION_ID_NONE;
    {   // the goal's own chunk: TARGET_TILE field + destination LOS (nav.c:1819-1847)
        pfnav_field_req rq = base;
        rq.chunk_r = tgt_chunk_r; rq.chunk_c = tgt_chunk_c; rq.target_type = PFNAV_TARGET_TILE;
        rq.tile_r = tgt_tile_r; rq.tile_c = tgt_tile_c;
        if (!emit_flow(rq, tchunk)) { pfnav_set_error("pfnav_plan_goal: flow output too small"); return PFNAV_ERR_NOMEM; }
        if (*n_los >= max_los) { pfnav_set_error("pfnav_plan_goal: los output too small"); return PFNAV_ERR_NOMEM; }
        pfnav_los_req lq;
        memset(&lq, 0, sizeof(lq));
        lq.chunk_r = tgt_chunk_r; lq.chunk_c = tgt_chunk_c; lq.layer = layer; lq.faction_id = PFNAV_FACTION_ID_NONE;
        lq.tgt_chunk_r = tgt_chunk_r; lq.tgt_chunk_c = tgt_chunk_c; lq.tgt_tile_r = tgt_tile_r; lq.tgt_tile_c = tgt_tile_c;
        lq.prev_index = -1;
        los_index[tchunk] = *n_los;
        los_out[*n_los] = lq; los_chunk[*n_los] = tchunk; (*n_los)++;
    }
    if (tli == 0xFFFF) return PFNAV_OK;      // goal tile impassable/blocked: nothing can flow to it
    seen[tchunk].push_back(tli);
    q.push_back({tchunk, tli});
    while (!q.empty()) {
        const node cur = q.front();
        q.pop_front();
        const auto &ports = ctx->portals[layer][cur.chunk];
        for (size_t pi = 0; pi < ports.size(); pi++) {
            const auto &P = ports[pi];                                // portal of the current (parent) chunk
            const auto &Q = ctx->portals[layer][P.conn_chunk][P.conn_idx];   // facing portal in the neighbour chunk
            const int nchunk = P.conn_chunk;
            // every island M of the neighbour that touches island cur.li across this portal
            const int len = (P.r1 - P.r0) + (P.c1 - P.c0) + 1;
            for (int k = 0; k < len; k++) {
                const int pr = P.r0 + (P.r1 > P.r0 ? k : 0), pc = P.c0 + (P.c1 > P.c0 ? k : 0);
                const int qr = Q.r0 + (Q.r1 > Q.r0 ? k : 0), qc = Q.c0 + (Q.c1 > Q.c0 ? k : 0);
                if (liid[(size_t)cur.chunk * 4096 + pr * 64 + pc] != cur.li) continue;
                const uint16_t M = liid[(size_t)nchunk * 4096 + qr * 64 + qc];
                if (M == 0xFFFF) continue;
                auto &sv = seen[nchunk];
                if (std::find(sv.begin(), sv.end(), M) != sv.end()) continue;
                sv.push_back(M);
                pfnav_field_req rq = base;
                rq.chunk_r = Q.chunk_r; rq.chunk_c = Q.chunk_c; rq.target_type = PFNAV_TARGET_PORTAL;
                rq.port_r0 = Q.r0; rq.port_c0 = Q.c0; rq.port_r1 = Q.r1; rq.port_c1 = Q.c1;
                rq.next_r0 = P.r0; rq.next_c0 = P.c0; rq.next_r1 = P.r1; rq.next_c1 = P.c1;
                rq.next_chunk_r = P.chunk_r; rq.next_chunk_c = P.chunk_c;
                rq.port_iid = M; rq.next_iid = cur.li;
                const bool first = nodes_in_chunk[nchunk] == 0;
                if (!emit_flow(rq, nchunk)) { pfnav_set_error("pfnav_plan_goal: flow output too small"); return PFNAV_ERR_NOMEM; }
                if (first) {
                    if (*n_los >= max_los) { pfnav_set_error("pfnav_plan_goal: los output too small"); return PFNAV_ERR_NOMEM; }
                    pfnav_los_req lq;
                    memset(&lq, 0, sizeof(lq));
                    lq.chunk_r = Q.chunk_r; lq.chunk_c = Q.chunk_c; lq.layer = layer; lq.faction_id = PFNAV_FACTION_ID_NONE;
                    lq.tgt_chunk_r = tgt_chunk_r; lq.tgt_chunk_c = tgt_chunk_c; lq.tgt_tile_r = tgt_tile_r; lq.tgt_tile_c = tgt_tile_c;
                    lq.prev_index = los_index[cur.chunk];
                    lq.prev_chunk_r = P.chunk_r; lq.prev_chunk_c = P.chunk_c;
                    los_index[nchunk] = *n_los;
                    los_out[*n_los] = lq; los_chunk[*n_los] = nchunk; (*n_los)++;
                }
                q.push_back({nchunk, M});
            }
        }
    }
    return PFNAV_OK;
}

// ------------------------------------------------------------------------------------------
// Pool slots with LRU eviction. The reference keeps its fields in LRU caches of fixed capacity
// (fieldcache.c:59-71; CONFIG_FLOW_CACHE_SZ / CONFIG_LOS_CACHE_SZ, config.h:64-65): a request that finds the
// cache full evicts the least recently used entry, and an agent that later misses it re-requests its path
// (nav.c:3484-3506 == pfnav_pool_repair). Here one slot holds the flow AND the LOS field of a (dest, chunk);
// "recently used" = the last tick whose desired-velocity pass read the slot (device side) or whose request
// named it (host side); ties go to the lower slot index.
// pf_pool_reserve: slots for the (dest, chunk) keys a request batch is about to write -- all or nothing: on
// PFNAV_ERR_NOMEM nothing was changed. keys may repeat. *out_evicted: slot-table entries of OTHER keys changed
// (the caller re-uploads the whole table and the `has` bytes).
// ------------------------------------------------------------------------------------------
int pf_pool_reserve(pfnav_ctx *ctx, const size_t *keys, size_t n, int32_t *slots_out, bool *out_evicted)
{
    if (out_evicted) *out_evicted = false;
    std::vector<size_t> fresh;
    for (size_t i = 0; i < n; i++)
        if (ctx->h_pool_slot[keys[i]] < 0) fresh.push_back(keys[i]);
    std::sort(fresh.begin(), fresh.end());
    fresh.erase(std::unique(fresh.begin(), fresh.end()), fresh.end());
    const size_t avail = ctx->pool_free.size() + (size_t)(ctx->pool_max - ctx->pool_used);
    if (fresh.size() > avail) {
        const size_t k = fresh.size() - avail;
        std::vector<uint32_t> stamp(ctx->h_slot_touch);
        if (ctx->device >= 0 && ctx->d_pool_touch) {
            PF_CUDA(cudaSetDevice(ctx->device));
            PF_CUDA(cudaDeviceSynchronize());         // ticks on any stream may still be stamping slots (eviction is rare)
            std::vector<uint32_t> dev(ctx->pool_max);
            PF_CUDA(cudaMemcpy(dev.data(), ctx->d_pool_touch, (size_t)ctx->pool_max * 4, cudaMemcpyDeviceToHost));
            for (int s = 0; s < ctx->pool_max; s++) stamp[s] = std::max(stamp[s], dev[s]);
        }
        std::vector<uint8_t> pinned(ctx->pool_max, 0);
        for (size_t i = 0; i < n; i++) { const int s = ctx->h_pool_slot[keys[i]]; if (s >= 0) pinned[s] = 1; }
        std::vector<int32_t> cand;
        for (int s = 0; s < ctx->pool_used; s++)
            if (ctx->h_slot_owner[s] >= 0 && !pinned[s]) cand.push_back(s);
        if (cand.size() < k) {
            pfnav_set_error("field pool full: %zu new (dest, chunk) fields requested, %zu slots free, %zu evictable of %d",
                            fresh.size(), avail, cand.size(), ctx->pool_max);
            return PFNAV_ERR_NOMEM;
        }
        std::sort(cand.begin(), cand.end(), [&](int32_t a, int32_t b) { return stamp[a] != stamp[b] ? stamp[a] < stamp[b] : a < b; });
        for (size_t i = 0; i < k; i++) {
            const int32_t s = cand[i];
            const size_t owner = (size_t)ctx->h_slot_owner[s];
            ctx->h_pool_slot[owner] = -1; ctx->h_pool_ffid[owner] = 0;
            ctx->h_pool_has[s] = 0; ctx->h_slot_owner[s] = -1;
            ctx->pool_free.push_back(s);
            ctx->pool_evictions++;
        }
        if (out_evicted) *out_evicted = true;
        ctx->goal_batch.valid = false;
    }
    for (size_t key : fresh) {
        int32_t s;
        if (!ctx->pool_free.empty()) { s = ctx->pool_free.back(); ctx->pool_free.pop_back(); }
        else s = ctx->pool_used++;
        ctx->h_pool_slot[key] = s; ctx->h_slot_owner[s] = (int64_t)key; ctx->h_pool_has[s] = 0;
    }
    for (size_t i = 0; i < n; i++) {
        const int32_t s = ctx->h_pool_slot[keys[i]];
        ctx->h_slot_touch[s] = ctx->tick_no;
        if (slots_out) slots_out[i] = s;
    }
    return PFNAV_OK;
}

// The flow waves and the LOS dependency chains of a goal batch each get a context-owned stream (the flow kernels
// need ~100 KB of shared memory per CTA and would otherwise queue the caller's stream behind the persistent LOS
// CTAs that hold most of it).
// The LOS dependency chains of a goal batch are latency-bound (one thread per field replays the
// reference's heap) and read nothing the flow kernels write: they run on ctx->field_stream, forked
// after everything already queued on the caller's stream and joined (pf_fields_join) by whichever
// entry point next reads or writes LOS fields of the pool. The flow kernels, the position-index
// rebuild and the cohesion pass overlap them.
static int los_fork(pfnav_ctx *ctx, cudaStream_t st)
{
    // duration of the previous batch (fork -> done), if it has finished: input of the two-phase policy
    if (ctx->los_inflight && cudaEventQuery(ctx->ev_los) == cudaSuccess) {
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, ctx->ev_fork, ctx->ev_los) == cudaSuccess) ctx->last_los_ms = ms;
    }
    cudaGetLastError();
    PF_CUDA(cudaEventRecord(ctx->ev_fork, st));
    PF_CUDA(cudaStreamWaitEvent(ctx->field_stream, ctx->ev_fork, 0));
    PF_CUDA(cudaStreamWaitEvent(ctx->flow_stream, ctx->ev_fork, 0));
    return 0;
}
static int los_forked(pfnav_ctx *ctx)
{
    PF_CUDA(cudaEventRecord(ctx->ev_flow, ctx->flow_stream));
    PF_CUDA(cudaEventRecord(ctx->ev_los, ctx->field_stream));
    ctx->los_inflight = true;
    return 0;
}

// n_request_path's field-building half, resident on the device, for a batch of goals: plan every
// goal, then run the flow waves and the LOS dependency waves of ALL goals together straight into the
// field pool (no host round trip of field bytes; one launch per wave, not per goal).
// targets: 4 ints per goal {chunk_r, chunk_c, tile_r, tile_c}. Asynchronous on `stream`.
extern "C" int pfnav_pool_request_goals(pfnav_ctx *ctx, int ngoals, const int32_t *dests, int layer,
                                        const int32_t *targets, void *stream, int *out_n_flow, int *out_n_los)
{
    return pfnav_pool_request_goals_ex(ctx, ngoals, dests, layer, targets, 0, stream, out_n_flow, out_n_los);
}

extern "C" int pfnav_pool_request_goals_ex(pfnav_ctx *ctx, int ngoals, const int32_t *dests, int layer,
                                           const int32_t *targets, uint32_t flags, void *stream, int *out_n_flow, int *out_n_los)
{
    const bool missing_only = (flags & PFNAV_REQUEST_MISSING_ONLY) != 0;
    PF_ARG(ctx && ctx->d_pool_slot, "pool not created");
    PF_NEED_DEVICE(ctx);
    PF_ARG(ngoals >= 0 && (ngoals == 0 || (dests && targets)), "goals");
    if (out_n_flow) *out_n_flow = 0;
    if (out_n_los) *out_n_los = 0;
    if (ngoals == 0) return PFNAV_OK;
    const int chunks = ctx->chunk_w * ctx->chunk_h;
    {   // fast path: the same batch on an unchanged map and pool -> the plan is still resident on the device
        auto &gb = ctx->goal_batch;
        if (!missing_only && gb.valid && gb.epoch == ctx->map_epoch && gb.layer == layer && (int)gb.dests.size() == ngoals &&
            memcmp(gb.dests.data(), dests, (size_t)ngoals * 4) == 0 && memcmp(gb.targets.data(), targets, (size_t)ngoals * 16) == 0) {
            PF_CUDA(cudaSetDevice(ctx->device));
            cudaStream_t st = pf_stream(ctx, stream);
            const uint8_t *dev = (const uint8_t *)ctx->d_plan_buf;
            int rc = los_fork(ctx, st);          // the LOS chains do not read flow fields: fork before the flow waves
            if (rc) return rc;
            for (size_t w = 0; w + 1 < gb.fwave_off.size(); w++) {
                const int first = gb.fwave_off[w], cnt = gb.fwave_off[w + 1] - first;
                if (cnt <= 0) continue;
                rc = pfnav_flow_launch(ctx, (const pfnav_field_req *)dev + first, cnt, ctx->d_pool_flow,
                                       (const int32_t *)(dev + gb.b_fr) + first, ctx->flow_stream);
                if (rc) return rc;
            }
            const int32_t one_wave[2] = {0, gb.nl};
            rc = pfnav_los_launch(ctx, (const pfnav_los_req *)(dev + gb.b_fr + gb.b_fs), gb.nl, ctx->d_pool_los,
                                  (const int32_t *)(dev + gb.b_fr + gb.b_fs + gb.b_lr), 1, one_wave, ctx->field_stream);
            if (rc) return rc;
            rc = los_forked(ctx);
            if (rc) return rc;
            if (out_n_flow) *out_n_flow = gb.nf;
            if (out_n_los) *out_n_los = gb.nl;
            return PFNAV_OK;
        }
        gb.valid = false;
    }
    PF_CUDA(pf_fields_sync(ctx));       // the plan buffer below may still be read by a forked LOS kernel
    const int cap = chunks * 8 + 8;
    std::vector<pfnav_field_req> all_fr;
    std::vector<pfnav_los_req> all_lr;
    std::vector<int32_t> all_fs, all_fw, all_ls, all_ld;
    std::vector<size_t> fkeys, lkeys;
    for (int g = 0; g < ngoals; g++) PF_ARG(dests[g] >= 0 && dests[g] < ctx->pool_ndests, "dest");
    // the goals are planned independently of each other (read-only walks over the host mirrors): one host thread each
    struct goal_plan { std::vector<pfnav_field_req> fr; std::vector<pfnav_los_req> lr; std::vector<int32_t> fc, fw, lc; int nf = 0, nl = 0, rc = 0; std::string err; };
    std::vector<goal_plan> plans(ngoals);
    {
        auto plan_one = [&](int g) {
            goal_plan &P = plans[g];
            P.fr.resize(cap); P.lr.resize(cap); P.fc.resize(cap); P.fw.resize(cap); P.lc.resize(cap);
            P.rc = pfnav_plan_goal(ctx, layer, targets[4 * g], targets[4 * g + 1], targets[4 * g + 2], targets[4 * g + 3],
                                   P.fr.data(), P.fc.data(), P.fw.data(), cap, &P.nf, P.lr.data(), P.lc.data(), cap, &P.nl);
            if (P.rc) P.err = pfnav_last_error();          // the error text is thread-local
        };
        const int nthreads = std::max(1, std::min({ngoals, (int)std::thread::hardware_concurrency(), 32}));
        if (nthreads == 1) { for (int g = 0; g < ngoals; g++) plan_one(g); }
        else {
            std::atomic<int> next{0};
            std::vector<std::thread> pool;
            auto worker = [&]() { for (int g = next.fetch_add(1); g < ngoals; g = next.fetch_add(1)) plan_one(g); };
            for (int t = 1; t < nthreads; t++) pool.emplace_back(worker);
            worker();
            for (auto &t : pool) t.join();
        }
    }
    for (int g = 0; g < ngoals; g++) {
        goal_plan &P = plans[g];
        if (P.rc) { pfnav_set_error("%s", P.err.c_str()); return P.rc; }
        const int nf = P.nf, nl = P.nl;
        const pfnav_field_req *fr = P.fr.data(); const pfnav_los_req *lr = P.lr.data();
        const int32_t *fc = P.fc.data(), *fw = P.fw.data(), *lc = P.lc.data();
        // PFNAV_REQUEST_MISSING_ONLY: fields the pool still holds for this destination are kept, like a field-cache hit
        // (fieldcache.c:138-141); a rebuilt LOS field whose parent chunk is kept reads the parent out of its pool slot
        auto held = [&](int chunk, uint8_t bit) {
            if (!missing_only) return false;
            const int sl = ctx->h_pool_slot[(size_t)dests[g] * chunks + chunk];
            return sl >= 0 && (ctx->h_pool_has[sl] & bit);
        };
        for (int i = 0; i < nf; i++) {
            if (held(fc[i], 1)) continue;
            all_fr.push_back(fr[i]); all_fw.push_back(fw[i]);
            fkeys.push_back((size_t)dests[g] * chunks + fc[i]);
        }
        std::vector<int> depth(nl, 0), newpos(nl, -1);
        for (int i = 0; i < nl; i++) {
            if (held(lc[i], 2)) continue;
            pfnav_los_req q = lr[i];
            if (q.prev_index >= 0) {
                if (newpos[q.prev_index] >= 0) { depth[i] = depth[q.prev_index] + 1; q.prev_index = newpos[q.prev_index]; }
                else {
                    q.prev_index = -2;      // parent kept: its absolute pool slot rides in _pad
                    q._pad = ctx->h_pool_slot[(size_t)dests[g] * chunks + q.prev_chunk_r * ctx->chunk_w + q.prev_chunk_c];
                }
            }
            newpos[i] = (int)all_lr.size();
            all_lr.push_back(q); all_ld.push_back(depth[i]);
            lkeys.push_back((size_t)dests[g] * chunks + lc[i]);
        }
    }
    if (all_fr.empty() && all_lr.empty()) return PFNAV_OK;
    const int nf = (int)all_fr.size(), nl = (int)all_lr.size();
    {   // slots for everything the batch writes, all or nothing (nothing is published before this succeeds)
        std::vector<size_t> keys(fkeys);
        keys.insert(keys.end(), lkeys.begin(), lkeys.end());
        std::vector<int32_t> slots(keys.size());
        int rc = pf_pool_reserve(ctx, keys.data(), keys.size(), slots.data(), nullptr);
        if (rc) return rc;
        all_fs.assign(slots.begin(), slots.begin() + nf);
        all_ls.assign(slots.begin() + nf, slots.end());
        for (int i = 0; i < nf; i++) { ctx->h_pool_has[all_fs[i]] |= 1; ctx->h_pool_req[all_fs[i]] = all_fr[i]; }
        for (int i = 0; i < nl; i++) ctx->h_pool_has[all_ls[i]] |= 2;
    }
    // stable order by wave / depth
    int maxw = 0, maxd = 0;
    for (int i = 0; i < nf; i++) maxw = std::max(maxw, all_fw[i]);
    for (int i = 0; i < nl; i++) maxd = std::max(maxd, all_ld[i]);
    std::vector<int32_t> fwave_off(maxw + 2, 0), lwave_off(maxd + 2, 0);
    for (int i = 0; i < nf; i++) fwave_off[all_fw[i] + 1]++;
    for (int w = 0; w <= maxw; w++) fwave_off[w + 1] += fwave_off[w];
    for (int i = 0; i < nl; i++) lwave_off[all_ld[i] + 1]++;
    for (int d = 0; d <= maxd; d++) lwave_off[d + 1] += lwave_off[d];
    std::vector<int> lnew(nl);
    const size_t b_fr = (size_t)nf * sizeof(pfnav_field_req), b_fs = (size_t)nf * 4;
    const size_t b_lr = (size_t)nl * sizeof(pfnav_los_req), b_ls = (size_t)nl * 4;
    const size_t total = b_fr + b_fs + b_lr + b_ls;
    std::vector<uint8_t> host(total);
    pfnav_field_req *hfr = (pfnav_field_req *)host.data();
    int32_t *hfs = (int32_t *)(host.data() + b_fr);
    pfnav_los_req *hlr = (pfnav_los_req *)(host.data() + b_fr + b_fs);
    int32_t *hls = (int32_t *)(host.data() + b_fr + b_fs + b_lr);
    {
        std::vector<int32_t> cur(fwave_off.begin(), fwave_off.end() - 1);
        for (int i = 0; i < nf; i++) { const int k = cur[all_fw[i]]++; hfr[k] = all_fr[i]; hfs[k] = all_fs[i]; }
        std::vector<int32_t> cur2(lwave_off.begin(), lwave_off.end() - 1);
        for (int i = 0; i < nl; i++) lnew[i] = cur2[all_ld[i]]++;
        for (int i = 0; i < nl; i++) {
            pfnav_los_req q = all_lr[i];
            if (q.prev_index >= 0) q.prev_index = lnew[q.prev_index];
            hlr[lnew[i]] = q; hls[lnew[i]] = all_ls[i];
        }
    }
    PF_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = pf_stream(ctx, stream);
    // one staging buffer per context: goal batches are serialised on one stream by the caller
    if (ctx->plan_buf_bytes < total) {
        PF_CUDA(cudaStreamSynchronize(st));
        cudaFree(ctx->d_plan_buf);
        ctx->d_plan_buf = nullptr; ctx->plan_buf_bytes = 0;
        PF_CUDA(cudaMalloc(&ctx->d_plan_buf, total * 2));
        ctx->plan_buf_bytes = total * 2;
    }
    uint8_t *dev = (uint8_t *)ctx->d_plan_buf;
    PF_CUDA(cudaMemcpyAsync(dev, host.data(), total, cudaMemcpyHostToDevice, st));
    PF_CUDA(cudaMemcpyAsync(ctx->d_pool_slot, ctx->h_pool_slot.data(), ctx->h_pool_slot.size() * 4, cudaMemcpyHostToDevice, st));
    PF_CUDA(cudaMemcpyAsync(ctx->d_pool_los + (size_t)ctx->pool_max * 4096, ctx->h_pool_has.data(), ctx->pool_max,
                            cudaMemcpyHostToDevice, st));
    PF_CUDA(cudaStreamSynchronize(st));      // `host` is pageable and about to go out of scope
    const pfnav_field_req *dfr = (const pfnav_field_req *)dev;
    const int32_t *dfs = (const int32_t *)(dev + b_fr);
    const pfnav_los_req *dlr = (const pfnav_los_req *)(dev + b_fr + b_fs);
    const int32_t *dls = (const int32_t *)(dev + b_fr + b_fs + b_lr);
    int rc = los_fork(ctx, st);
    if (rc) return rc;
    for (int w = 0; w <= maxw; w++) {
        const int first = fwave_off[w], cnt = fwave_off[w + 1] - first;
        if (cnt <= 0) continue;
        rc = pfnav_flow_launch(ctx, dfr + first, cnt, ctx->d_pool_flow, dfs + first, ctx->flow_stream);
        if (rc) return rc;
    }
    if (nl) rc = pfnav_los_launch(ctx, dlr, nl, ctx->d_pool_los, dls, maxd + 1, lwave_off.data(), ctx->field_stream);
    if (rc) return rc;
    rc = los_forked(ctx);
    if (rc) return rc;
    {
        auto &gb = ctx->goal_batch;
        gb.valid = !missing_only; gb.epoch = ctx->map_epoch; gb.layer = layer;
        gb.dests.assign(dests, dests + ngoals); gb.targets.assign(targets, targets + 4 * (size_t)ngoals);
        gb.nf = nf; gb.nl = nl; gb.fwave_off = fwave_off; gb.b_fr = b_fr; gb.b_fs = b_fs; gb.b_lr = b_lr; gb.b_ls = b_ls;
    }
    if (out_n_flow) *out_n_flow = nf;
    if (out_n_los) *out_n_los = nl;
    return PFNAV_OK;
}

extern "C" int pfnav_pool_request_goal(pfnav_ctx *ctx, int dest, int layer, int tgt_chunk_r, int tgt_chunk_c,
                                       int tgt_tile_r, int tgt_tile_c, void *stream, int *out_n_flow, int *out_n_los)
{
    const int32_t d = dest, t[4] = {tgt_chunk_r, tgt_chunk_c, tgt_tile_r, tgt_tile_c};
    return pfnav_pool_request_goals(ctx, 1, &d, layer, t, stream, out_n_flow, out_n_los);
}

// Order `stream` after the LOS chains that pfnav_pool_request_goals forked onto the context's field
// stream. pfnav_agents_tick and the pool entry points do this themselves.
extern "C" int pfnav_fields_join(pfnav_ctx *ctx, void *stream)
{
    PF_ARG(ctx, "ctx");
    PF_NEED_DEVICE(ctx);
    PF_CUDA(cudaSetDevice(ctx->device));
    PF_CUDA(pf_fields_join(ctx, pf_stream(ctx, stream)));
    return PFNAV_OK;
}
