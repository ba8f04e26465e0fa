// pfnav_route.cu -- host-side, cost-faithful restatement of the reference's hierarchical path
// request (the part of the hot path that decides WHICH chunk fields a destination needs).
//
// Restates (reference file:line):
//   n_update_island_field / n_visit_island              src/navigation/nav.c:1731-1772, 856-901
//   n_link_chunk_portals + AStar_GridPath               src/navigation/nav.c:593-636, a_star.c:303-427
//   n_build_portal_travel_index + N_GridNeighbours      src/navigation/nav.c:1314-1363, 4878-4910
//   n_update_edge_states / n_local_ports_connected      src/navigation/nav.c:668-717
//   N_PortalReachableFromTile                           src/navigation/nav.c:4852
//   N_ClosestPathableLocalIsland                        src/navigation/nav.c:5131
//   n_closest_reachable_portal / _from_location         src/navigation/nav.c:1365, 1398
//   AStar_PortalGraphPath + neighbours_portal_graph     src/navigation/a_star.c:429-552, 212-262
//   n_request_path                                      src/navigation/nav.c:1774-2047
//   N_FlowFieldID                                       src/navigation/field.c:1952
//
// Everything here is pointer-heavy host logic over a tiny graph; the fields it asks for are built by
// the CUDA kernels. Float costs follow the reference's evaluation order exactly (same heap, same
// neighbour order), because the chosen portal path depends on them.
#include "pfnav_internal.cuh"
#include <algorithm>
#include <atomic>
#include <thread>
#include <deque>
#include <float.h>
#include <math.h>
#include <string.h>
#include <unordered_map>

namespace {

struct coord { int r, c; };
struct tdesc { int chunk_r, chunk_c, tile_r, tile_c; };

// pqueue.h:109-208 (1-indexed heap, float priority, hole sift)
template <class T>
struct pq {
    struct node { float prio; T data; };
    std::vector<node> n{1};
    int size = 0;
    void push(float prio, const T &d)
    {
        if ((int)n.size() < size + 2) n.resize(std::max<size_t>(32, n.size() * 2));
        int curr = size + 1, parent = curr / 2;
        while (curr > 1 && n[parent].prio > prio) { n[curr] = n[parent]; curr = parent; parent /= 2; }
        n[curr].prio = prio; n[curr].data = d;
        size++;
    }
    T pop()
    {
        T out = n[1].data;
        n[1] = n[size--];
        int root = 1;
        while (root != size + 1) {
            int target = size + 1;
            const int l = root * 2, r = l + 1;
            if (l <= size && n[l].prio < n[target].prio) target = l;
            if (r <= size && n[r].prio < n[target].prio) target = r;
            n[root] = n[target];
            root = target;
        }
        return out;
    }
};

// N_GridNeighbours (nav.c:4878) == a_star.c neighbours_grid (:101)
static int grid_neighbours(const uint8_t *cost, coord at, coord *out, float *costs)
{
    int ret = 0;
    for (int r = -1; r <= 1; r++)
        for (int c = -1; c <= 1; c++) {
            const int ar = at.r + r, ac = at.c + c;
            if (ar < 0 || ar >= 64 || ac < 0 || ac >= 64) continue;
            if (r == 0 && c == 0) continue;
            if (cost[ar * 64 + ac] == 0xFF) continue;
            const bool diag = (r == c) || (r == -c);
            if (diag && cost[ar * 64 + at.c] == 0xFF && cost[at.r * 64 + ac] == 0xFF) continue;
            const float cost_mult = diag ? (float)sqrt(2) : 1.0f;
            out[ret] = {ar, ac};
            costs[ret] = cost[ar * 64 + ac] * cost_mult;
            ret++;
        }
    return ret;
}

// heuristic (a_star.c:282)
static float heuristic(coord a, coord b)
{
    const float D = 1.0f;
    const float D2 = sqrt(2) * D;
    const int dx = abs(a.r - b.r), dy = abs(a.c - b.c);
    return D * (dx + dy) + (D2 - 2 * D) * std::min(dx, dy);
}

// AStar_GridPath (a_star.c:303), cost only
static bool astar_grid_cost(const uint8_t *cost, coord start, coord finish, float *out_cost)
{
    pq<coord> frontier;
    float running[4096];
    bool has_cost[4096], has_from[4096];
    memset(has_cost, 0, sizeof(has_cost));
    memset(has_from, 0, sizeof(has_from));
    running[start.r * 64 + start.c] = 0.0f;
    has_cost[start.r * 64 + start.c] = true;
    frontier.push(0.0f, start);
    while (frontier.size > 0) {
        const coord curr = frontier.pop();
        if (curr.r == finish.r && curr.c == finish.c) break;
        coord nb[8]; float nc[8];
        const int n = grid_neighbours(cost, curr, nb, nc);
        for (int i = 0; i < n; i++) {
            const int k = nb[i].r * 64 + nb[i].c;
            const float new_cost = running[curr.r * 64 + curr.c] + nc[i];
            if (!has_cost[k] || new_cost < running[k]) {
                running[k] = new_cost; has_cost[k] = true;
                frontier.push(new_cost + heuristic(finish, nb[i]), nb[i]);
                has_from[k] = true;
            }
        }
    }
    if (!has_from[finish.r * 64 + finish.c]) return false;
    *out_cost = running[finish.r * 64 + finish.c];
    return true;
}

static inline uint16_t cost_pack(float cost)          // portal_cost_pack (nav_data.h:56)
{
    if (cost == FLT_MAX) return 0xffff;
    const float scaled = cost * 8 + 0.5f;
    if (scaled >= 0xffff) return 0xffff - 1;
    return (uint16_t)scaled;
}
static inline float cost_unpack(uint16_t p) { return p == 0xffff ? FLT_MAX : (float)p / 8; }

}   // namespace

struct pfnav_route_edge { int es; int nb; float cost; };          // nb: portal index in the same chunk
struct pfnav_route_chunk {
    std::vector<std::vector<pfnav_route_edge>> edges;             // per portal
    std::vector<uint16_t> travel;                                 // [nportals][4096]
};
struct pfnav_route_layer {
    bool built = false;
    uint64_t generation = 0;                                      // bumped by every pfnav_route_build: device copies follow it
    std::vector<pfnav_route_chunk> chunks;
    std::vector<uint16_t> islands;                                // global islands, chunk-blocked
};
// The routing structures live in the context (pfnav_ctx::route_state): two contexts driven from two threads
// (one per GPU in one process) share nothing.
typedef std::vector<pfnav_route_layer> route_vec;
static inline route_vec *routes_of(const pfnav_ctx *ctx) { return (route_vec *)ctx->route_state; }
// the built layer, or nullptr
static inline pfnav_route_layer *route_layer(const pfnav_ctx *ctx, int layer)
{
    route_vec *rv = routes_of(ctx);
    if (!rv || layer < 0 || layer >= (int)rv->size() || !(*rv)[layer].built) return nullptr;
    return &(*rv)[layer];
}

void pfnav_route_dev_forget(pfnav_ctx *ctx);
void pfnav_route_forget(pfnav_ctx *ctx) { pfnav_route_dev_forget(ctx); delete routes_of(ctx); ctx->route_state = nullptr; }
// the portal lists of `layer` were rebuilt: its edge / travel tables no longer describe them
void pfnav_route_invalidate_layer(pfnav_ctx *ctx, int layer)
{
    route_vec *rv = routes_of(ctx);
    if (rv && layer >= 0 && layer < (int)rv->size()) (*rv)[layer] = pfnav_route_layer();
}

static inline const uint8_t *L_cost(const pfnav_ctx *ctx, int layer, int chunk)
{ return ctx->h_cost.data() + ((size_t)layer * ctx->chunk_w * ctx->chunk_h + chunk) * 4096; }
static inline const uint16_t *L_blk(const pfnav_ctx *ctx, int layer, int chunk)
{ return ctx->h_blk.data() + ((size_t)layer * ctx->chunk_w * ctx->chunk_h + chunk) * 4096; }
static inline const uint16_t *L_liid(const pfnav_ctx *ctx, int layer, int chunk)
{ return ctx->h_liid.data() + ((size_t)layer * ctx->chunk_w * ctx->chunk_h + chunk) * 4096; }

// n_update_edge_states (nav.c:693) for one chunk
static void update_edge_states(pfnav_ctx *ctx, pfnav_route_layer &RL, int layer, int chunk)
{
    const auto &ports = ctx->portals[layer][chunk];
    const uint16_t *blk = L_blk(ctx, layer, chunk), *li = L_liid(ctx, layer, chunk);
    for (size_t i = 0; i < ports.size(); i++)
        for (auto &e : RL.chunks[chunk].edges[i]) {
            const auto &a = ports[i], &b = ports[e.nb];
            bool conn = false;
            for (int r1 = a.r0; r1 <= a.r1 && !conn; r1++)
                for (int c1 = a.c0; c1 <= a.c1 && !conn; c1++) {
                    if (blk[r1 * 64 + c1] > 0) continue;
                    for (int r2 = b.r0; r2 <= b.r1 && !conn; r2++)
                        for (int c2 = b.c0; c2 <= b.c1; c2++) {
                            if (blk[r2 * 64 + c2] > 0) continue;
                            if (li[r1 * 64 + c1] == li[r2 * 64 + c2]) { conn = true; break; }
                        }
                }
            e.es = conn ? 0 : 1;       // EDGE_STATE_ACTIVE : EDGE_STATE_BLOCKED
        }
}

// n_update_edge_states for one chunk after its islands changed; returns the number of flipped edges,
// or -1 when the routing structure of the layer has not been built.
int pfnav_route_refresh_edges(pfnav_ctx *ctx, int layer, int chunk)
{
    pfnav_route_layer *prl = route_layer(ctx, layer);
    if (!prl) return -1;
    auto &RL = *prl;
    std::vector<int> before;
    for (auto &ev : RL.chunks[chunk].edges) for (auto &e : ev) before.push_back(e.es);
    update_edge_states(ctx, RL, layer, chunk);
    int flipped = 0; size_t k = 0;
    for (auto &ev : RL.chunks[chunk].edges) for (auto &e : ev) flipped += (e.es != before[k++]);
    return flipped;
}

// Build the routing structure of one layer (needs pfnav_map_build_nav first).
extern "C" int pfnav_route_build(pfnav_ctx *ctx, int layer)
{
    PF_ARG(ctx && ctx->d_cost, "map not created");
    PF_ARG(layer >= 0 && layer < ctx->nlayers, "layer");
    PF_ARG((size_t)layer < ctx->portals.size() && !ctx->portals[layer].empty(), "pfnav_map_build_nav not called for this layer");
    if (!ctx->route_state) ctx->route_state = new route_vec();
    auto &RV = *routes_of(ctx);
    if ((int)RV.size() < ctx->nlayers) RV.resize(ctx->nlayers);
    pfnav_route_layer &RL = RV[layer];
    const int cw = ctx->chunk_w, chh = ctx->chunk_h, chunks = cw * chh;
    RL.chunks.assign(chunks, {});
    // every chunk's portal edges and travel index depend on that chunk alone: spread the chunks over the host cores
    auto build_chunk = [&](int ch) {
        const auto &ports = ctx->portals[layer][ch];
        const uint8_t *cost = L_cost(ctx, layer, ch);
        auto &RC = RL.chunks[ch];
        RC.edges.assign(ports.size(), {});
        // n_link_chunk_portals (nav.c:593): A* between portal centres, i-major j-minor
        for (size_t i = 0; i < ports.size(); i++)
            for (size_t j = 0; j < ports.size(); j++) {
                if (i == j) continue;
                const coord a = {(ports[i].r0 + ports[i].r1) / 2, (ports[i].c0 + ports[i].c1) / 2};
                const coord b = {(ports[j].r0 + ports[j].r1) / 2, (ports[j].c0 + ports[j].c1) / 2};
                float c;
                if (astar_grid_cost(cost, a, b, &c)) RC.edges[i].push_back({0, (int)j, c});
            }
        // n_build_portal_travel_index (nav.c:1314): FIFO expansion, first visit fixes the cost
        RC.travel.assign(ports.size() * 4096, 0xffff);
        for (size_t pi = 0; pi < ports.size(); pi++) {
            bool visited[4096];
            memset(visited, 0, sizeof(visited));
            std::deque<std::pair<float, coord>> q;
            for (int r = ports[pi].r0; r <= ports[pi].r1; r++)
                for (int c = ports[pi].c0; c <= ports[pi].c1; c++) { q.push_back({0.0f, {r, c}}); visited[r * 64 + c] = true; }
            while (!q.empty()) {
                const auto cur = q.front();
                q.pop_front();
                RC.travel[pi * 4096 + cur.second.r * 64 + cur.second.c] = cost_pack(cur.first);
                coord nb[8]; float nc[8];
                const int n = grid_neighbours(cost, cur.second, nb, nc);
                for (int i = 0; i < n; i++) {
                    const int k = nb[i].r * 64 + nb[i].c;
                    if (visited[k]) continue;
                    q.push_back({cur.first + nc[i], nb[i]});
                    visited[k] = true;
                }
            }
        }
    };
    {
        const int nthreads = std::max(1, std::min({(int)std::thread::hardware_concurrency(), 32, chunks}));
        std::atomic<int> next{0};
        std::vector<std::thread> pool;
        for (int t = 1; t < nthreads; t++)
            pool.emplace_back([&]() { for (int ch = next.fetch_add(1); ch < chunks; ch = next.fetch_add(1)) build_chunk(ch); });
        for (int ch = next.fetch_add(1); ch < chunks; ch = next.fetch_add(1)) build_chunk(ch);
        for (auto &t : pool) t.join();
    }
    // n_update_island_field (nav.c:1731): global islands ignoring blockers, ids from 0
    RL.islands.assign((size_t)chunks * 4096, 0xffff);
    {
        uint16_t id = 0;
        std::deque<tdesc> q;
        for (int cr = 0; cr < chh; cr++)
            for (int cc = 0; cc < cw; cc++)
                for (int t = 0; t < 4096; t++) {
                    const int ch = cr * cw + cc;
                    if (RL.islands[(size_t)ch * 4096 + t] != 0xffff || L_cost(ctx, layer, ch)[t] == 0xFF) continue;
                    RL.islands[(size_t)ch * 4096 + t] = id;
                    q.push_back({cr, cc, t >> 6, t & 63});
                    while (!q.empty()) {
                        const tdesc cur = q.front();
                        q.pop_front();
                        const int dr[4] = {0, 0, -1, 1}, dc[4] = {-1, 1, 0, 0};
                        for (int e = 0; e < 4; e++) {
                            const int ar = cur.chunk_r * 64 + cur.tile_r + dr[e], ac = cur.chunk_c * 64 + cur.tile_c + dc[e];
                            if (ar < 0 || ar >= chh * 64 || ac < 0 || ac >= cw * 64) continue;
                            const int nch = (ar >> 6) * cw + (ac >> 6), nt = (ar & 63) * 64 + (ac & 63);
                            if (RL.islands[(size_t)nch * 4096 + nt] == 0xffff && L_cost(ctx, layer, nch)[nt] != 0xFF) {
                                RL.islands[(size_t)nch * 4096 + nt] = id;
                                q.push_back({ar >> 6, ac >> 6, ar & 63, ac & 63});
                            }
                        }
                    }
                    id++;
                }
    }
    for (int ch = 0; ch < chunks; ch++) update_edge_states(ctx, RL, layer, ch);
    { static std::atomic<uint64_t> next_generation{1}; RL.generation = next_generation++; }
    RL.built = true;
    return PFNAV_OK;
}

// Read-back for parity tests: islands and the edge table of one portal (neighbour ref, state, cost)
extern "C" int pfnav_route_islands_get(pfnav_ctx *ctx, int layer, uint16_t *out)
{
    PF_ARG(ctx && out, "args");
    const pfnav_route_layer *prl = route_layer(ctx, layer);
    PF_ARG(prl, "pfnav_route_build not called");
    memcpy(out, prl->islands.data(), prl->islands.size() * 2);
    return PFNAV_OK;
}

extern "C" int pfnav_route_edges_get(pfnav_ctx *ctx, int layer, int chunk, int portal, uint32_t *out, int maxout, int *out_n)
{
    PF_ARG(ctx && out && out_n, "args");
    const pfnav_route_layer *prl = route_layer(ctx, layer);
    PF_ARG(prl, "pfnav_route_build not called");
    const auto &RL = *prl;
    PF_ARG(chunk >= 0 && chunk < (int)RL.chunks.size() && portal >= 0 && portal < (int)RL.chunks[chunk].edges.size(), "chunk/portal");
    int n = 0;
    for (const auto &e : RL.chunks[chunk].edges[portal]) {
        if (n >= maxout) break;
        out[n * 3 + 0] = ((uint32_t)chunk << 8) | (uint32_t)e.nb;        // portal_ref_make (nav_data.h:90)
        out[n * 3 + 1] = (uint32_t)e.es;
        memcpy(&out[n * 3 + 2], &e.cost, 4);
        n++;
    }
    *out_n = n;
    return PFNAV_OK;
}

namespace {

struct Router {
    pfnav_ctx *ctx; pfnav_route_layer &RL; int layer; int cw, chh;
    const pfnav_ctx::portal_t &P(int chunk, int idx) const { return ctx->portals[layer][chunk][idx]; }
    int nports(int chunk) const { return (int)ctx->portals[layer][chunk].size(); }
    const uint16_t *li(int chunk) const { return L_liid(ctx, layer, chunk); }

    // N_PortalReachableFromTile (nav.c:4852)
    bool portal_reachable_from_tile(int chunk, int pi, coord tile) const
    {
        const auto &p = P(chunk, pi);
        const uint16_t *l = li(chunk);
        for (int r = p.r0; r <= p.r1; r++)
            for (int c = p.c0; c <= p.c1; c++) {
                if (l[r * 64 + c] == 0xffff) continue;
                for (int r1 = tile.r - 1; r1 <= tile.r + 1; r1++)
                    for (int c1 = tile.c - 1; c1 <= tile.c + 1; c1++) {
                        if (r1 < 0 || r1 >= 64 || c1 < 0 || c1 >= 64) continue;
                        if (l[r * 64 + c] == l[r1 * 64 + c1]) return true;
                    }
            }
        return false;
    }

    // N_ClosestPathableLocalIsland (nav.c:5131): note `visited[tile_r * tile_h + tile_c]`
    uint16_t closest_pathable_liid(int chunk, coord target) const
    {
        const uint16_t *l = li(chunk);
        if (l[target.r * 64 + target.c] != 0xffff) return l[target.r * 64 + target.c];
        bool visited[4096];
        memset(visited, 0, sizeof(visited));
        std::deque<coord> q;
        q.push_back(target);
        visited[target.r * 64 + target.c] = true;
        while (!q.empty()) {
            const coord cur = q.front();
            q.pop_front();
            const int dr[4] = {0, 0, -1, 1}, dc[4] = {-1, 1, 0, 0};
            for (int e = 0; e < 4; e++) {
                const int nr = cur.r + dr[e], nc = cur.c + dc[e];
                // M_Tile_RelativeDesc succeeds across chunk borders; tiles outside this chunk are skipped
                if (nr < 0 || nr >= 64 || nc < 0 || nc >= 64) continue;
                if (visited[nr * 64 + nc]) continue;
                if (l[nr * 64 + nc] != 0xffff) return l[nr * 64 + nc];
                visited[nr * 64 + nc] = true;
                q.push_back({nr, nc});
            }
        }
        return 0xffff;
    }

    // n_closest_reachable_portal (nav.c:1365)
    int closest_reachable_portal(int chunk, coord start, bool unblocked) const
    {
        int ret = -1;
        float min_cost = FLT_MAX;
        for (int i = 0; i < nports(chunk); i++) {
            const float cost = cost_unpack(RL.chunks[chunk].travel[(size_t)i * 4096 + start.r * 64 + start.c]);
            if (unblocked && !portal_reachable_from_tile(chunk, i, start)) continue;
            if (cost < min_cost) { ret = i; min_cost = cost; }
        }
        return ret;
    }

    struct hop { int chunk, pi; uint16_t liid; };
    static uint64_t hop_key(const hop &h) { return ((uint64_t)h.liid << 32) | ((uint64_t)h.chunk << 8) | (uint64_t)h.pi; }

    // neighbours_portal_graph (a_star.c:212) incl. portal_reachable_from_island (:138) and
    // portal_connected_liids (:151)
    int neighbours(const hop &cur, hop *out, float *costs) const
    {
        int ret = 0;
        const int maxout = 256;
        const uint16_t *l = li(cur.chunk);
        for (const auto &e : RL.chunks[cur.chunk].edges[cur.pi]) {
            if (ret == maxout) return ret;
            if (e.es == 1) continue;
            const auto &np = P(cur.chunk, e.nb);
            bool reach = false;
            for (int r = np.r0; r <= np.r1 && !reach; r++)
                for (int c = np.c0; c <= np.c1; c++)
                    if (l[r * 64 + c] == cur.liid) { reach = true; break; }
            if (!reach) continue;
            out[ret] = {cur.chunk, e.nb, cur.liid};
            costs[ret] = e.cost;
            ret++;
        }
        const auto &p = P(cur.chunk, cur.pi);
        const auto &conn = P(p.conn_chunk, p.conn_idx);
        const uint16_t *cl = li(p.conn_chunk);
        uint16_t conn_liids[256];
        int nconn = 0;
        for (int r1 = p.r0; r1 <= p.r1; r1++)
            for (int c1 = p.c0; c1 <= p.c1; c1++) {
                if (l[r1 * 64 + c1] != cur.liid) continue;
                // the reference scans every tile of the connected portal for Manhattan distance 1 (a_star.c:151);
                // only the four neighbours of (r1, c1) can qualify: visit those inside the portal's rectangle in the
                // same (row, column) order
                const int ar = p.chunk_r * 64 + r1, ac = p.chunk_c * 64 + c1;
                const int cand[4][2] = {{ar - 1, ac}, {ar, ac - 1}, {ar, ac + 1}, {ar + 1, ac}};
                for (int k = 0; k < 4; k++) {
                    const int r2 = cand[k][0] - conn.chunk_r * 64, c2 = cand[k][1] - conn.chunk_c * 64;
                    if (r2 < conn.r0 || r2 > conn.r1 || c2 < conn.c0 || c2 > conn.c1) continue;
                    if (nconn == 256) goto done_conn;
                    const uint16_t nl = cl[r2 * 64 + c2];
                    bool contains = false;
                    for (int i = 0; i < nconn; i++) if (conn_liids[i] == nl) { contains = true; break; }
                    if (!contains && nl != 0xffff) conn_liids[nconn++] = nl;
                }
            }
    done_conn:
        for (int i = 0; i < nconn; i++) {
            if (ret == maxout) return ret;
            out[ret] = {p.conn_chunk, p.conn_idx, conn_liids[i]};
            costs[ret] = 1;
            ret++;
        }
        return ret;
    }

    // AStar_PortalGraphPath (a_star.c:429)
    bool portal_graph_path(tdesc start, tdesc end, int fin_chunk, int fin_pi, std::vector<hop> &path, float *out_cost) const
    {
        const int bchunk = start.chunk_r * cw + start.chunk_c, echunk = end.chunk_r * cw + end.chunk_c;
        const uint16_t start_liid = closest_pathable_liid(bchunk, {start.tile_r, start.tile_c});
        if (start_liid == 0xffff) return false;
        const uint16_t end_liid = closest_pathable_liid(echunk, {end.tile_r, end.tile_c});
        if (end_liid == 0xffff) return false;
        pq<hop> frontier;
        std::unordered_map<uint64_t, float> running;
        std::unordered_map<uint64_t, hop> came_from;
        for (int i = 0; i < nports(bchunk); i++) {
            const coord tc = {start.tile_r, start.tile_c};
            if (!portal_reachable_from_tile(bchunk, i, tc)) continue;
            const float cost = cost_unpack(RL.chunks[bchunk].travel[(size_t)i * 4096 + tc.r * 64 + tc.c]);
            if (cost != FLT_MAX) {
                const hop h = {bchunk, i, start_liid};
                running[hop_key(h)] = cost;
                frontier.push(cost, h);
            }
        }
        const float penalty = (float)sqrt(pow(64, 2.0f) + pow(64, 2.0f));     // portal_node_penalty (a_star.c:298)
        while (frontier.size > 0) {
            const hop cur = frontier.pop();
            if (cur.chunk == fin_chunk && cur.pi == fin_pi && cur.liid == end_liid) break;
            hop nb[256]; float nc[256];
            const int n = neighbours(cur, nb, nc);
            for (int i = 0; i < n; i++) {
                const float new_cost = running[hop_key(cur)] + nc[i] + penalty;
                auto it = running.find(hop_key(nb[i]));
                if (it == running.end() || new_cost < it->second) {
                    running[hop_key(nb[i])] = new_cost;
                    frontier.push(new_cost, nb[i]);
                    came_from[hop_key(nb[i])] = cur;
                }
            }
        }
        const hop last = {fin_chunk, fin_pi, end_liid};
        if (came_from.find(hop_key(last)) == came_from.end()) return false;
        path.clear();
        hop cur = last;
        while (true) {
            path.push_back(cur);
            auto it = came_from.find(hop_key(cur));
            if (it == came_from.end()) break;
            cur = it->second;
        }
        std::reverse(path.begin(), path.end());
        *out_cost = running[hop_key(last)];
        return true;
    }

    // n_closest_reachable_from_location (nav.c:1398)
    int closest_reachable_from_location(int chunk, tdesc loc, tdesc *out_nearest) const
    {
        float shortest = FLT_MAX;
        int ret = -1;
        tdesc nearest = {0, 0, 0, 0};
        std::vector<hop> path;
        const uint16_t *l = li(chunk);
        for (int i = 0; i < nports(chunk); i++) {
            uint16_t liids[64];
            int nl = 0;
            const auto &p = P(chunk, i);
            for (int r = p.r0; r <= p.r1; r++)
                for (int c = p.c0; c <= p.c1; c++) {
                    const tdesc cur = {p.chunk_r, p.chunk_c, r, c};
                    const uint16_t cl = l[r * 64 + c];
                    bool contains = false;
                    for (int k = 0; k < nl; k++) if (liids[k] == cl) { contains = true; break; }
                    if (!contains && nl < 64) {
                        liids[nl++] = cl;
                        float cost;
                        if (portal_graph_path(loc, cur, chunk, i, path, &cost) && cost < shortest) {
                            nearest = cur; shortest = cost; ret = i;
                        }
                    }
                }
        }
        if (ret >= 0) *out_nearest = nearest;
        return ret;
    }

    bool blocked_off(int chunk, coord tile) const       // n_blocked_off (nav.c:1544)
    {
        for (int i = 0; i < nports(chunk); i++) if (portal_reachable_from_tile(chunk, i, tile)) return false;
        return true;
    }
    bool normally_reachable(int chunk, coord a, coord b) const      // n_normally_reachable (nav.c:1531)
    {
        for (int i = 0; i < nports(chunk); i++) {
            const bool ar = RL.chunks[chunk].travel[(size_t)i * 4096 + a.r * 64 + a.c] != 0xffff;
            const bool br = RL.chunks[chunk].travel[(size_t)i * 4096 + b.r * 64 + b.c] != 0xffff;
            if (ar != br) return false;
        }
        return true;
    }
};

// M_Tile_DescForPoint2D with the nav resolution (tile.c:547)
static bool desc_for_point(const pfnav_ctx *ctx, float px, float pz, tdesc *out)
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

}   // namespace

// n_request_path (nav.c:1774-2047), request-generation half. Emits, in the reference's order, the
// flow requests (with the N_FlowFieldID of each, field.c:1952) and LOS requests a cold field cache
// would have to build for (src -> dst), given which (dest, chunk) entries `have_flow` / `have_los`
// already hold (arrays of `chunks` entries; have_flow holds the ff_id mapped for the chunk or 0).
// Returns 1 in *out_ok when a path exists (the function's bool), plus the dest_id.
extern "C" int pfnav_route_request_path(pfnav_ctx *ctx, int layer, float src_x, float src_z, float dst_x, float dst_z,
                                        const uint64_t *have_flow, const uint8_t *have_los,
                                        pfnav_field_req *flow_out, uint64_t *flow_ffid, int32_t *flow_chunk, int max_flow,
                                        int *n_flow, pfnav_los_req *los_out, int32_t *los_chunk, int max_los, int *n_los,
                                        uint32_t *out_dest_id, int *out_ok)
{
    PF_ARG(ctx && n_flow && n_los && out_ok && out_dest_id, "args");
    pfnav_route_layer *prl = route_layer(ctx, layer);
    PF_ARG(prl, "pfnav_route_build not called");
    pfnav_route_layer &RL = *prl;
    const int cw = ctx->chunk_w, chunks = cw * ctx->chunk_h;
    Router R{ctx, RL, layer, cw, ctx->chunk_h};
    *n_flow = 0; *n_los = 0; *out_ok = 0;
    // n_update_dirty_local_islands + n_update_all_edge_states (nav.c:1786-1787)
    for (int ch = 0; ch < chunks; ch++) update_edge_states(ctx, RL, layer, ch);
    tdesc src, dst;
    PF_ARG(desc_for_point(ctx, src_x, src_z, &src) && desc_for_point(ctx, dst_x, dst_z, &dst), "position outside the map");
    const uint32_t dest_id = (((uint32_t)dst.chunk_r & 0x3f) << 26) | (((uint32_t)dst.chunk_c & 0x3f) << 20) |
                             (((uint32_t)dst.tile_r & 0x3f) << 14) | (((uint32_t)dst.tile_c & 0x3f) << 8) |
                             (((uint32_t)layer & 0xf) << 4) | ((uint32_t)ctx->req_faction & 0xfu);   // n_dest_id (nav.c:839)
    *out_dest_id = dest_id;
    const int schunk = src.chunk_r * cw + src.chunk_c, dchunk = dst.chunk_r * cw + dst.chunk_c;
    if (RL.islands[(size_t)schunk * 4096 + src.tile_r * 64 + src.tile_c] != RL.islands[(size_t)dchunk * 4096 + dst.tile_r * 64 + dst.tile_c])
        return PFNAV_OK;
    std::vector<uint64_t> mapped(have_flow, have_flow + chunks);      // local copy of the (dest, chunk) -> ffid mapping
    std::vector<uint8_t> los_have(have_los, have_los + chunks);
    std::vector<int> los_index(chunks, -1);
    auto ffid_tile = [&](int chunk, int tr, int tc) -> uint64_t {
        return ((uint64_t)layer << 60) | ((uint64_t)1 << 56) | ((uint64_t)tr << 24) | ((uint64_t)tc << 16) |
               ((uint64_t)(chunk / cw) << 8) | (uint64_t)(chunk % cw);
    };
    auto emit_flow = [&](const pfnav_field_req &q, uint64_t id, int chunk) -> bool {
        if (*n_flow >= max_flow) return false;
        flow_out[*n_flow] = q; flow_ffid[*n_flow] = id; flow_chunk[*n_flow] = chunk; (*n_flow)++;
        return true;
    };
    auto emit_los = [&](int chunk, int prev_chunk) -> bool {
        if (*n_los >= max_los) return false;
        pfnav_los_req q;
        memset(&q, 0, sizeof(q));
        q.chunk_r = chunk / cw; q.chunk_c = chunk % cw; q.layer = layer; q.faction_id = ctx->req_faction;
        q.tgt_chunk_r = dst.chunk_r; q.tgt_chunk_c = dst.chunk_c; q.tgt_tile_r = dst.tile_r; q.tgt_tile_c = dst.tile_c;
        // -1: destination chunk; >= 0: request of this batch; -2: the previous chunk's field already
        // exists (have_los) and the executor substitutes its pool slot into _pad
        q.prev_index = prev_chunk < 0 ? -1 : (los_index[prev_chunk] >= 0 ? los_index[prev_chunk] : -2);
        if (prev_chunk >= 0) { q.prev_chunk_r = prev_chunk / cw; q.prev_chunk_c = prev_chunk % cw; }
        los_index[chunk] = *n_los;
        los_out[*n_los] = q; los_chunk[*n_los] = chunk; (*n_los)++;
        los_have[chunk] = 1;
        return true;
    };
    pfnav_field_req base;
    memset(&base, 0, sizeof(base));
    base.layer = layer; base.faction_id = ctx->req_faction;
    // destination chunk field + LOS (nav.c:1815-1847)
    if (!mapped[dchunk]) {
        pfnav_field_req q = base;
        q.chunk_r = dst.chunk_r; q.chunk_c = dst.chunk_c; q.target_type = PFNAV_TARGET_TILE; q.init = 1;
        q.tile_r = dst.tile_r; q.tile_c = dst.tile_c;
        const uint64_t id = ffid_tile(dchunk, dst.tile_r, dst.tile_c);
        if (!emit_flow(q, id, dchunk)) { pfnav_set_error("flow output too small"); return PFNAV_ERR_NOMEM; }
        mapped[dchunk] = id;
    }
    if (!los_have[dchunk] && !emit_los(dchunk, -1)) { pfnav_set_error("los output too small"); return PFNAV_ERR_NOMEM; }
    const uint16_t dst_cl = R.closest_pathable_liid(dchunk, {dst.tile_r, dst.tile_c});
    if (schunk == dchunk && R.closest_pathable_liid(schunk, {src.tile_r, src.tile_c}) == dst_cl) { *out_ok = 1; return PFNAV_OK; }
    if (schunk == dchunk) {
        const bool either = R.blocked_off(schunk, {src.tile_r, src.tile_c}) || R.blocked_off(schunk, {dst.tile_r, dst.tile_c});
        if (either && R.normally_reachable(schunk, {src.tile_r, src.tile_c}, {dst.tile_r, dst.tile_c})) { *out_ok = 1; return PFNAV_OK; }
    }
    int dst_port = R.closest_reachable_portal(dchunk, {dst.tile_r, dst.tile_c}, true);
    if (dst_port < 0) dst_port = R.closest_reachable_portal(dchunk, {dst.tile_r, dst.tile_c}, false);
    if (dst_port < 0) return PFNAV_OK;
    std::vector<Router::hop> path;
    float cost;
    bool exists = R.portal_graph_path(src, dst, dchunk, dst_port, path, &cost);
    if (!exists) {
        const tdesc orig = dst;
        tdesc nd = dst;
        dst_port = R.closest_reachable_from_location(dchunk, src, &nd);
        if (dst_port >= 0) dst = nd;
        if (R.closest_pathable_liid(dchunk, {dst.tile_r, dst.tile_c}) != R.closest_pathable_liid(dchunk, {orig.tile_r, orig.tile_c}) &&
            src.chunk_r == dst.chunk_r && src.chunk_c == dst.chunk_c) { *out_ok = 1; return PFNAV_OK; }
        if (dst_port >= 0) exists = R.portal_graph_path(src, dst, dchunk, dst_port, path, &cost);
    }
    if (!exists) {
        if (src.chunk_r == dst.chunk_r && src.chunk_c == dst.chunk_c) *out_ok = 1;
        return PFNAV_OK;
    }
    int prev_los_chunk = dst.chunk_r * cw + dst.chunk_c;
    const uint16_t dst_liid_now = R.closest_pathable_liid(dchunk, {dst.tile_r, dst.tile_c});
    // walk the portal path backwards (nav.c:1941-2042)
    for (int i = (int)path.size() - 1; i > 0; i--) {
        int next_hop_idx = i;
        if (i == 1 && path[i].chunk != schunk) next_hop_idx = 0;
        const Router::hop curr_hop = path[std::max(next_hop_idx - 1, 0)];
        const Router::hop next_hop = path[next_hop_idx];
        const auto &cp = R.P(curr_hop.chunk, curr_hop.pi);
        if (cp.conn_chunk == next_hop.chunk && cp.conn_idx == next_hop.pi) continue;
        if (curr_hop.chunk == dchunk && next_hop.chunk == dchunk && next_hop.pi == dst_port && next_hop.liid == dst_liid_now) continue;
        const int chunk = curr_hop.chunk;
        const auto &np = R.P(next_hop.chunk, next_hop.pi);
        const auto &nn = R.P(np.conn_chunk, np.conn_idx);
        pfnav_field_req q = base;
        q.chunk_r = chunk / cw; q.chunk_c = chunk % cw; q.target_type = PFNAV_TARGET_PORTAL;
        q.port_r0 = np.r0; q.port_c0 = np.c0; q.port_r1 = np.r1; q.port_c1 = np.c1;
        q.next_r0 = nn.r0; q.next_c0 = nn.c0; q.next_r1 = nn.r1; q.next_c1 = nn.c1;
        q.next_chunk_r = nn.chunk_r; q.next_chunk_c = nn.chunk_c;
        q.port_iid = next_hop.liid;
        q.next_iid = (i < (int)path.size() - 1) ? path[next_hop_idx + 1].liid : dst_liid_now;
        // N_FlowFieldID TARGET_PORTAL (field.c:1954)
        const uint64_t new_id = ((uint64_t)layer << 60) | ((uint64_t)0 << 56) | (((uint64_t)q.next_iid & 0xf) << 48) |
                                (((uint64_t)q.port_iid & 0xf) << 40) | ((uint64_t)np.r0 << 34) | ((uint64_t)np.c0 << 28) |
                                ((uint64_t)np.r1 << 22) | ((uint64_t)np.c1 << 16) | ((uint64_t)q.chunk_r << 8) | (uint64_t)q.chunk_c;
        if (mapped[chunk]) {
            if (mapped[chunk] != new_id) {
                q.init = 0;          // merge into the chunk's existing field (nav.c:1994-2010)
                if (!emit_flow(q, new_id, chunk)) { pfnav_set_error("flow output too small"); return PFNAV_ERR_NOMEM; }
                mapped[chunk] = new_id;
            }
        } else {
            q.init = 1;
            if (!emit_flow(q, new_id, chunk)) { pfnav_set_error("flow output too small"); return PFNAV_ERR_NOMEM; }
            mapped[chunk] = new_id;
        }
        if (!los_have[chunk]) {
            if (!emit_los(chunk, prev_los_chunk)) { pfnav_set_error("los output too small"); return PFNAV_ERR_NOMEM; }
        }
        prev_los_chunk = chunk;
    }
    *out_ok = 1;
    return PFNAV_OK;
}

// N_RequestPath (nav.c:3386) against the device field pool: route src -> dst exactly as the reference
// does, then build only the fields the pool does not hold yet, in place, on the device.
extern "C" int pfnav_pool_request_path(pfnav_ctx *ctx, int dest, int layer, float src_x, float src_z, float dst_x,
                                       float dst_z, void *stream, uint32_t *out_dest_id, int *out_ok, int *out_n_flow,
                                       int *out_n_los)
{
    PF_ARG(ctx && ctx->d_pool_slot, "pool not created");
    PF_NEED_DEVICE(ctx);
    PF_ARG(dest >= 0 && dest < ctx->pool_ndests, "dest");
    PF_CUDA(cudaSetDevice(ctx->device));
    PF_CUDA(pf_fields_sync(ctx));
    const int chunks = ctx->chunk_w * ctx->chunk_h;
    std::vector<uint8_t> have_los(chunks, 0);
    for (int c = 0; c < chunks; c++) {
        const int s = ctx->h_pool_slot[(size_t)dest * chunks + c];
        have_los[c] = (s >= 0 && (ctx->h_pool_has[s] & 2)) ? 1 : 0;
    }
    const int cap = chunks * 4 + 8;
    std::vector<pfnav_field_req> fr(cap);
    std::vector<pfnav_los_req> lr(cap);
    std::vector<uint64_t> fid(cap);
    std::vector<int32_t> fc(cap), lc(cap);
    int nf = 0, nl = 0, ok = 0;
    uint32_t did = 0;
    int rc = pfnav_route_request_path(ctx, layer, src_x, src_z, dst_x, dst_z, ctx->h_pool_ffid.data() + (size_t)dest * chunks,
                                      have_los.data(), fr.data(), fid.data(), fc.data(), cap, &nf, lr.data(), lc.data(), cap,
                                      &nl, &did, &ok);
    if (rc) return rc;
    if (out_dest_id) *out_dest_id = did;
    if (out_ok) *out_ok = ok;
    if (out_n_flow) *out_n_flow = nf;
    if (out_n_los) *out_n_los = nl;
    if (nf == 0 && nl == 0) return PFNAV_OK;
    ctx->goal_batch.valid = false;          // the staging buffer is about to be reused
    // flow waves: a request that updates a chunk already written in this batch runs one wave later
    std::vector<int32_t> fslot(nf), fwave(nf), lslot(nl), ldepth(nl, 0);
    std::vector<int> seen(chunks, 0);
    int maxw = 0, maxd = 0;
    bool evicted = false;
    {   // slots for everything this request writes (and the previous-chunk LOS fields it reads), all or nothing
        std::vector<size_t> keys;
        for (int i = 0; i < nf; i++) keys.push_back((size_t)dest * chunks + fc[i]);
        for (int i = 0; i < nl; i++) keys.push_back((size_t)dest * chunks + lc[i]);
        for (int i = 0; i < nl; i++)
            if (lr[i].prev_index == -2) keys.push_back((size_t)dest * chunks + lr[i].prev_chunk_r * ctx->chunk_w + lr[i].prev_chunk_c);
        std::vector<int32_t> slots(keys.size());
        rc = pf_pool_reserve(ctx, keys.data(), keys.size(), slots.data(), &evicted);
        if (rc) return rc;
        for (int i = 0; i < nf; i++) { fslot[i] = slots[i]; ctx->h_pool_has[fslot[i]] |= 1; }
        for (int i = 0; i < nl; i++) { lslot[i] = slots[nf + i]; ctx->h_pool_has[lslot[i]] |= 2; }
    }
    for (int i = 0; i < nf; i++) {
        fwave[i] = seen[fc[i]]++;
        maxw = std::max(maxw, fwave[i]);
        ctx->h_pool_ffid[(size_t)dest * chunks + fc[i]] = fid[i];
        ctx->h_pool_req[fslot[i]] = fr[i];
    }
    for (int i = 0; i < nl; i++) {
        if (lr[i].prev_index >= 0) ldepth[i] = ldepth[lr[i].prev_index] + 1;
        else if (lr[i].prev_index == -2) {
            const int pchunk = lr[i].prev_chunk_r * ctx->chunk_w + lr[i].prev_chunk_c;
            lr[i]._pad = ctx->h_pool_slot[(size_t)dest * chunks + pchunk];       // absolute pool slot of prev_los
        }
        maxd = std::max(maxd, ldepth[i]);
    }
    std::vector<int32_t> fwave_off(maxw + 2, 0), lwave_off(maxd + 2, 0);
    for (int i = 0; i < nf; i++) fwave_off[fwave[i] + 1]++;
    for (int w = 0; w <= maxw; w++) fwave_off[w + 1] += fwave_off[w];
    for (int i = 0; i < nl; i++) lwave_off[ldepth[i] + 1]++;
    for (int d = 0; d <= maxd; d++) lwave_off[d + 1] += lwave_off[d];
    const size_t b_fr = (size_t)nf * sizeof(pfnav_field_req), b_fs = (size_t)nf * 4;
    const size_t b_lr = (size_t)nl * sizeof(pfnav_los_req), b_ls = (size_t)nl * 4;
    const size_t total = b_fr + b_fs + b_lr + b_ls;
    std::vector<uint8_t> host(total ? total : 1);
    pfnav_field_req *hfr = (pfnav_field_req *)host.data();
    int32_t *hfs = (int32_t *)(host.data() + b_fr);
    pfnav_los_req *hlr = (pfnav_los_req *)(host.data() + b_fr + b_fs);
    int32_t *hls = (int32_t *)(host.data() + b_fr + b_fs + b_lr);
    {
        std::vector<int32_t> cur(fwave_off.begin(), fwave_off.end() - 1), cur2(lwave_off.begin(), lwave_off.end() - 1), lnew(nl);
        for (int i = 0; i < nf; i++) { const int k = cur[fwave[i]]++; hfr[k] = fr[i]; hfs[k] = fslot[i]; }
        for (int i = 0; i < nl; i++) lnew[i] = cur2[ldepth[i]]++;
        for (int i = 0; i < nl; i++) {
            pfnav_los_req q = lr[i];
            if (q.prev_index >= 0) q.prev_index = lnew[q.prev_index];
            hlr[lnew[i]] = q; hls[lnew[i]] = lslot[i];
        }
    }
    PF_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = pf_stream(ctx, stream);
    if (ctx->plan_buf_bytes < total) {
        PF_CUDA(cudaStreamSynchronize(st));
        cudaFree(ctx->d_plan_buf);
        ctx->d_plan_buf = nullptr; ctx->plan_buf_bytes = 0;
        PF_CUDA(cudaMalloc(&ctx->d_plan_buf, total * 2));
        ctx->plan_buf_bytes = total * 2;
    }
    uint8_t *dev = (uint8_t *)ctx->d_plan_buf;
    PF_CUDA(cudaMemcpyAsync(dev, host.data(), total, cudaMemcpyHostToDevice, st));
    if (evicted)      // entries of other destinations lost their slots
        PF_CUDA(cudaMemcpyAsync(ctx->d_pool_slot, ctx->h_pool_slot.data(), ctx->h_pool_slot.size() * 4, cudaMemcpyHostToDevice, st));
    else
        PF_CUDA(cudaMemcpyAsync(ctx->d_pool_slot + (size_t)dest * chunks, ctx->h_pool_slot.data() + (size_t)dest * chunks,
                                (size_t)chunks * 4, cudaMemcpyHostToDevice, st));
    PF_CUDA(cudaMemcpyAsync(ctx->d_pool_los + (size_t)ctx->pool_max * 4096, ctx->h_pool_has.data(), ctx->pool_max,
                            cudaMemcpyHostToDevice, st));
    PF_CUDA(cudaStreamSynchronize(st));
    for (int w = 0; w <= maxw; w++) {
        const int first = fwave_off[w], cnt = fwave_off[w + 1] - first;
        if (cnt <= 0) continue;
        rc = pfnav_flow_launch(ctx, (const pfnav_field_req *)dev + first, cnt, ctx->d_pool_flow, (const int32_t *)(dev + b_fr) + first, st);
        if (rc) return rc;
    }
    if (nl) rc = pfnav_los_launch(ctx, (const pfnav_los_req *)(dev + b_fr + b_fs), nl, ctx->d_pool_los,
                                  (const int32_t *)(dev + b_fr + b_fs + b_lr), maxd + 1, lwave_off.data(), st);
    return rc;
}

// Read one pool entry back to the host (4096 B each; either pointer may be NULL). *out_has: bit0 flow, bit1 LOS.
extern "C" int pfnav_pool_get(pfnav_ctx *ctx, int dest, int chunk_r, int chunk_c, uint8_t *flow_out, uint8_t *los_out,
                              int *out_has, uint64_t *out_ffid)
{
    PF_ARG(ctx && ctx->d_pool_slot && out_has, "args");
    PF_NEED_DEVICE(ctx);
    PF_ARG(dest >= 0 && dest < ctx->pool_ndests && chunk_r >= 0 && chunk_r < ctx->chunk_h && chunk_c >= 0 && chunk_c < ctx->chunk_w, "dest/chunk");
    const size_t si = (size_t)dest * ctx->chunk_w * ctx->chunk_h + chunk_r * ctx->chunk_w + chunk_c;
    const int slot = ctx->h_pool_slot[si];
    *out_has = slot < 0 ? 0 : ctx->h_pool_has[slot];
    if (out_ffid) *out_ffid = ctx->h_pool_ffid[si];
    if (slot < 0) return PFNAV_OK;
    PF_CUDA(cudaSetDevice(ctx->device));
    PF_CUDA(cudaDeviceSynchronize());
    if (flow_out && (*out_has & 1)) PF_CUDA(cudaMemcpy(flow_out, ctx->d_pool_flow + (size_t)slot * 4096, 4096, cudaMemcpyDeviceToHost));
    if (los_out && (*out_has & 2)) PF_CUDA(cudaMemcpy(los_out, ctx->d_pool_los + (size_t)slot * 4096, 4096, cudaMemcpyDeviceToHost));
    return PFNAV_OK;
}

// ------------------------------------------------------------------------------------------
// arrived() support (movement.c:2170-2196): the two map searches it makes depend only on the flock
// target and the layer, not on the entity, so they are evaluated once per (flock, layer) on the host
// mirrors and handed to the state-update kernel as constants:
//   * N_ClosestPathable (nav.c:4126): nearest not-blocked tile to the target, breadth-first over the
//     4-neighbourhood in the order {0,-1} {0,+1} {-1,0} {+1,0}; result = M_Tile_Bounds corner (x, z)
//   * N_IsMaximallyClose (nav.c:4707) -> n_closest_island_tiles (nav.c:1226, ignore_blockers = false):
//     the tiles of the target's global island, without blockers, at the smallest Manhattan distance
//     from the target (the target tile itself is never reported: it is marked visited up front and the
//     {0,0} delta is skipped), as "tile centres" map_pos -/+ (abs tile index) * 4 (nav.c:4729-4732)
// ------------------------------------------------------------------------------------------
int pfnav_arrival_consts(pfnav_ctx *ctx, int layer, float tx, float tz, pf_arrival_consts *out)
{
    out->nearest_ok = 0; out->nearest[0] = out->nearest[1] = 0.0f; out->mc_n = 0;
    const pfnav_route_layer *prl = route_layer(ctx, layer);
    if (!prl) {
        pfnav_set_error("pfnav_agents_compute_updates: pfnav_route_build(layer %d) is needed (global islands, nav.c:1731)", layer);
        return PFNAV_ERR_ARG;
    }
    const pfnav_route_layer &RL = *prl;
    tdesc t;
    if (!desc_for_point(ctx, tx, tz, &t)) return PFNAV_OK;       // target outside the map: the reference asserts
    const int cw = ctx->chunk_w, chh = ctx->chunk_h, W = cw * 64, H = chh * 64;
    auto blocked = [&](int ar, int ac) {
        const int ch = (ar >> 6) * cw + (ac >> 6), tt = (ar & 63) * 64 + (ac & 63);
        return L_cost(ctx, layer, ch)[tt] == 0xFF || L_blk(ctx, layer, ch)[tt] > 0;
    };
    const int tr = t.chunk_r * 64 + t.tile_r, tc = t.chunk_c * 64 + t.tile_c;
    // ---- N_ClosestPathable ----
    if (!blocked(tr, tc)) {
        out->nearest_ok = 1; out->nearest[0] = tx; out->nearest[1] = tz;
    } else {
        std::vector<uint8_t> vis((size_t)W * H, 0);
        std::vector<int> q; q.push_back(tr * W + tc);
        for (size_t qi = 0; qi < q.size(); qi++) {
            const int ar = q[qi] / W, ac = q[qi] % W;
            if (!blocked(ar, ac)) {
                // M_Tile_Bounds (tile.c:356): x decreases with the column, z increases with the row
                out->nearest_ok = 1;
                out->nearest[0] = (ctx->map_x - (float)((ac >> 6) * 256)) - (float)((ac & 63) * 4);
                out->nearest[1] = (ctx->map_z + (float)((ar >> 6) * 256)) + (float)((ar & 63) * 4);
                break;
            }
            const int dr[4] = {0, 0, -1, 1}, dc[4] = {-1, 1, 0, 0};
            for (int e = 0; e < 4; e++) {
                const int nr = ar + dr[e], nc = ac + dc[e];
                if (nr < 0 || nr >= H || nc < 0 || nc >= W) continue;
                if (vis[(size_t)nr * W + nc]) continue;
                vis[(size_t)nr * W + nc] = 1;
                q.push_back(nr * W + nc);
            }
        }
    }
    // ---- n_closest_island_tiles(target, giid(target), ignore_blockers = false, maxout = 256) ----
    {
        const uint16_t giid = RL.islands[(size_t)(t.chunk_r * cw + t.chunk_c) * 4096 + t.tile_r * 64 + t.tile_c];
        std::vector<uint8_t> vis((size_t)W * H, 0);
        std::vector<int> q; q.push_back(tr * W + tc);
        vis[(size_t)tr * W + tc] = 1;
        int first_mh = -1, n = 0;
        bool done = false;
        for (size_t qi = 0; qi < q.size() && !done; qi++) {
            const int ar = q[qi] / W, ac = q[qi] % W;
            const int dr[9] = {0, 0, 0, -1, 1, -1, -1, 1, 1}, dc[9] = {0, -1, 1, 0, 0, -1, 1, -1, 1};
            for (int e = 0; e < 9; e++) {
                const int nr = ar + dr[e], nc = ac + dc[e];
                if (nr < 0 || nr >= H || nc < 0 || nc >= W) continue;
                if (vis[(size_t)nr * W + nc]) continue;
                const int ch = (nr >> 6) * cw + (nc >> 6), tt = (nr & 63) * 64 + (nc & 63);
                bool skip = RL.islands[(size_t)ch * 4096 + tt] != giid;
                if (L_blk(ctx, layer, ch)[tt] > 0) skip = true;
                const int mh = abs(tr - nr) + abs(tc - nc);
                if (first_mh > 0 && mh > first_mh) { done = true; break; }
                if (!skip) {
                    out->mc[n][0] = ctx->map_x - (float)nc * 4.0f;      // (chunk_c*64 + tile_c) * tile_dims.x
                    out->mc[n][1] = ctx->map_z + (float)nr * 4.0f;
                    n++;
                    if (first_mh == -1) first_mh = mh;
                    if (n == PF_ARRIVAL_MC_MAX) { done = true; break; }
                }
                vis[(size_t)nr * W + nc] = 1;
                q.push_back(nr * W + nc);
            }
        }
        out->mc_n = n;
    }
    return PFNAV_OK;
}

// ------------------------------------------------------------------------------------------
// Seeds of the repair chain of N_DesiredPointSeekVelocity (nav.c:3508-3554), host side (they need the
// global islands and breadth-first ring searches over the chunk; the integration itself runs on the
// device, k_flow_repair):
//   kind 0  N_FlowFieldUpdateToNearestPathable (field.c:2247): the passable tiles that border the
//           non-passable blob containing (start_r, start_c)   (field_passable_frontier, field.c:1441)
//   kind 1  N_FlowFieldUpdateIslandToNearest (field.c:2307): the tiles of local island `local_iid`
//           nearest (Manhattan) to the field's own frontier     (field_closest_tiles_local, field.c:1010)
// mask: 64 rows of 64 bits.
// ------------------------------------------------------------------------------------------
int pfnav_repair_seeds(pfnav_ctx *ctx, const pfnav_field_req &q, int kind, int arg, uint64_t *mask)
{
    memset(mask, 0, 64 * sizeof(uint64_t));
    const int layer = q.layer, cw = ctx->chunk_w;
    const int chunk = q.chunk_r * cw + q.chunk_c;
    const uint8_t *cost = L_cost(ctx, layer, chunk);
    const uint16_t *blk = L_blk(ctx, layer, chunk), *liid = L_liid(ctx, layer, chunk);
    auto passable = [&](int r, int c) { return cost[r * 64 + c] != 0xFF && blk[r * 64 + c] == 0; };
    static const int dr4[4] = {0, 0, -1, 1}, dc4[4] = {-1, 1, 0, 0};
    if (kind == 0) {
        const int sr = arg >> 8, sc = arg & 0xFF;
        PF_ARG(sr >= 0 && sr < 64 && sc >= 0 && sc < 64, "repair start tile");
        PF_ARG(!passable(sr, sc), "N_FlowFieldUpdateToNearestPathable needs a non-passable start tile (field.c:1455)");
        std::vector<uint8_t> vis(4096, 0);
        std::vector<int> fq; fq.push_back(sr * 64 + sc);
        vis[sr * 64 + sc] = 1;
        for (size_t qi = 0; qi < fq.size(); qi++) {
            const int r = fq[qi] >> 6, c = fq[qi] & 63;
            if (passable(r, c)) { mask[r] |= 1ull << c; continue; }
            for (int e = 0; e < 4; e++) {
                const int ar = r + dr4[e], ac = c + dc4[e];
                if (ar < 0 || ar >= 64 || ac < 0 || ac >= 64 || vis[ar * 64 + ac]) continue;
                vis[ar * 64 + ac] = 1;
                fq.push_back(ar * 64 + ac);
            }
        }
        return PFNAV_OK;
    }
    // ---- kind 1 ----
    const pfnav_route_layer *prl = route_layer(ctx, layer);
    PF_ARG(prl, "pfnav_route_build(layer) is needed (global islands, nav.c:1731)");
    const uint16_t *gisl = prl->islands.data() + (size_t)chunk * 4096;
    const uint16_t local_iid = (uint16_t)arg;
    std::vector<int> init;
    if ((q.target_type & 0xFF) >= 2) {
        // TARGET_ENEMIES / TARGET_ENTITY field of pool destination q._pad: the entities' own tiles inside the chunk
        int rc = pfnav_aux_chunk_seeds(ctx, q._pad, q.chunk_r, q.chunk_c, init);
        if (rc) return rc;
    } else if ((q.target_type & 0xFF) == PFNAV_TARGET_TILE) {
        // field_tile_initial_frontier (field.c:1096); when the tile is blocked the reference retries with
        // ignoreblock (field.c:2367) -- either way the frontier is the tile itself
        init.push_back(q.tile_r * 64 + q.tile_c);
    } else {
        const uint16_t *nliid = L_liid(ctx, layer, q.next_chunk_r * cw + q.next_chunk_c);
        for (int r = q.port_r0; r <= q.port_r1; r++)
            for (int c = q.port_c0; c <= q.port_c1; c++) {
                if (!passable(r, c)) continue;
                if (q.port_iid != PFNAV_ISLAND_NONE && liid[r * 64 + c] != q.port_iid) continue;
                bool adj = false;
                for (int r2 = q.next_r0; r2 <= q.next_r1 && !adj; r2++)
                    for (int c2 = q.next_c0; c2 <= q.next_c1; c2++) {
                        const int ddr = (q.next_chunk_r * 64 + r2) - (q.chunk_r * 64 + r);
                        const int ddc = (q.next_chunk_c * 64 + c2) - (q.chunk_c * 64 + c);
                        if (abs(ddr) + abs(ddc) == 1 && nliid[r2 * 64 + c2] == q.next_iid) { adj = true; break; }
                    }
                if (adj) init.push_back(r * 64 + c);
            }
    }
    int min_mh = INT_MAX;
    std::vector<int> newf;
    std::vector<uint8_t> vis(4096);
    std::vector<int> fq, tmp;
    for (int t0 : init) {
        const int r0 = t0 >> 6, c0 = t0 & 63;
        const uint16_t cg = gisl[t0], cl = liid[t0];
        if (cl == local_iid) {
            if (min_mh > 0) newf.clear();
            min_mh = 0;
            newf.push_back(t0);
            continue;
        }
        // field_closest_tiles_local(chunk, curr, local_iid, curr_giid)
        std::fill(vis.begin(), vis.end(), 0);
        fq.clear(); tmp.clear();
        fq.push_back(t0); vis[t0] = 1;
        int first = -1;
        const size_t cap = 4096 - newf.size();
        for (size_t qi = 0; qi < fq.size(); qi++) {
            const int r = fq[qi] >> 6, c = fq[qi] & 63;
            for (int e = 0; e < 4; e++) {
                const int ar = r + dr4[e], ac = c + dc4[e];
                if (ar < 0 || ar >= 64 || ac < 0 || ac >= 64 || vis[ar * 64 + ac]) continue;
                vis[ar * 64 + ac] = 1;
                fq.push_back(ar * 64 + ac);
            }
            const int mh = abs(r0 - r) + abs(c0 - c);
            if (first > -1 && mh > first) break;
            if (cost[r * 64 + c] == 0xFF) continue;
            if (blk[r * 64 + c] > 0) continue;
            if (cg != PFNAV_ISLAND_NONE && gisl[r * 64 + c] != cg) continue;
            if (local_iid != PFNAV_ISLAND_NONE && liid[r * 64 + c] != local_iid) continue;
            if (first == -1) first = mh;
            tmp.push_back(r * 64 + c);
            if (tmp.size() == cap) break;
        }
        if (tmp.empty()) continue;
        const int mh = abs((tmp[0] >> 6) - r0) + abs((tmp[0] & 63) - c0);
        if (mh < min_mh) { min_mh = mh; newf.clear(); }
        if (mh > min_mh) continue;
        newf.insert(newf.end(), tmp.begin(), tmp.end());
    }
    for (int t : newf) mask[t >> 6] |= 1ull << (t & 63);
    return PFNAV_OK;
}

// N_RequestPathAttacking (nav.c:3393): the faction carried by the path requests that follow -- packed into the
// dest_id and forwarded to every flow / LOS request, exactly what n_request_path does with its faction_id.
extern "C" int pfnav_request_faction(pfnav_ctx *ctx, int faction_id)
{
    PF_ARG(ctx, "ctx");
    PF_ARG(faction_id == PFNAV_FACTION_ID_NONE || (faction_id >= 0 && faction_id < 15), "faction_id");
    ctx->req_faction = faction_id;
    return PFNAV_OK;
}

// Inspection / test entry: the per-(target, layer) constants of arrived() (movement.c:2170) that
// pfnav_agents_compute_updates uploads: N_ClosestPathable's answer and the tile centres N_IsMaximallyClose
// compares with. out_mc: up to cap (x, z) pairs. Host structure code; needs pfnav_route_build(layer).
extern "C" int pfnav_route_arrival_consts(pfnav_ctx *ctx, int layer, float tx, float tz, int32_t *out_nearest_ok,
                                          float *out_nearest_xz, float *out_mc_xz, size_t cap, int32_t *out_mc_n)
{
    PF_ARG(ctx && out_nearest_ok && out_nearest_xz && out_mc_n, "null argument");
    pf_arrival_consts c;
    int rc = pfnav_arrival_consts(ctx, layer, tx, tz, &c);
    if (rc) return rc;
    *out_nearest_ok = c.nearest_ok; out_nearest_xz[0] = c.nearest[0]; out_nearest_xz[1] = c.nearest[1];
    *out_mc_n = c.mc_n;
    for (int i = 0; i < c.mc_n && (size_t)i < cap && out_mc_xz; i++) { out_mc_xz[2 * i] = c.mc[i][0]; out_mc_xz[2 * i + 1] = c.mc[i][1]; }
    return PFNAV_OK;
}

// ------------------------------------------------------------------------------------------
// Device-side AStar_PortalGraphPath (a_star.c:429): a batch of portal-graph searches, one thread per
// search. The search itself is the reference's, statement for statement -- N_ClosestPathableLocalIsland (nav.c:5131)
// for both ends, N_PortalReachableFromTile (:4852) + the travel-cost index for the start portals,
// neighbours_portal_graph (a_star.c:212), portal_node_penalty (:298), and the 1-indexed binary heap of pqueue.h:109-208
// whose tie order decides between equal-cost paths -- over device copies of the portal table, the intra-chunk edges with
// their states, the travel-cost index and the local-island image. Per-search scratch (heap, cost / predecessor table,
// flood queue) lives in HBM: 0.6 MB per search in flight, a few hundred searches per launch.
// Not wired into pfnav_pool_request_path yet: the rest of n_request_path (fallback branches, field emission) is host code.
// ------------------------------------------------------------------------------------------
namespace {

struct dev_portal { int16_t r0, c0, r1, c1, chunk_r, chunk_c; int32_t conn_chunk, conn_idx, edge_off, edge_cnt; };
struct dev_edge { int32_t nb, es; float cost; };
struct dev_route {
    const dev_portal *ports;      // [chunks][64]
    const int32_t *nports;        // [chunks]
    const int32_t *port_base;     // [chunks]: first row of the chunk's portals in `travel`
    const dev_edge *edges;
    const uint16_t *travel;       // [port_base[chunk] + i][4096]
    const uint16_t *liid;         // image of the layer [H64][W64]
    int cw, chh, W64;
    float penalty;
};
#define GP_HT 8192                // cost / predecessor table slots (open addressing)
#define GP_PQ 16384               // heap capacity
struct gp_scratch {
    uint64_t hkey[GP_HT]; uint64_t hfrom[GP_HT]; float hcost[GP_HT];
    float pprio[GP_PQ]; uint64_t pdata[GP_PQ];
    uint16_t queue[4096]; uint8_t visited[4096];
};
struct route_dev_state {
    dev_portal *d_ports = nullptr; int32_t *d_nports = nullptr, *d_port_base = nullptr; dev_edge *d_edges = nullptr;
    uint16_t *d_travel = nullptr; size_t travel_rows = 0, nedges = 0;
    gp_scratch *d_scratch = nullptr; int scratch_n = 0;
    void *d_req = nullptr, *d_out = nullptr; size_t req_bytes = 0, out_bytes = 0;
    int layer = -1; uint64_t epoch = ~0ull, generation = 0;
};

__device__ __forceinline__ uint16_t gp_li(const dev_route &R, int chunk, int r, int c)
{
    return R.liid[(size_t)((chunk / R.cw) * 64 + r) * R.W64 + (chunk % R.cw) * 64 + c];
}
__device__ __forceinline__ uint64_t gp_key(int chunk, int pi, uint16_t liid) { return ((uint64_t)liid << 32) | ((uint64_t)chunk << 8) | (uint64_t)pi; }

// N_ClosestPathableLocalIsland (nav.c:5131)
__device__ uint16_t gp_closest_liid(const dev_route &R, gp_scratch &S, int chunk, int tr, int tc)
{
    const uint16_t own = gp_li(R, chunk, tr, tc);
    if (own != 0xffff) return own;
    for (int i = 0; i < 4096; i++) S.visited[i] = 0;
    int head = 0, tail = 0;
    S.queue[tail++] = (uint16_t)(tr * 64 + tc);
    S.visited[tr * 64 + tc] = 1;
    while (head < tail) {
        const int cur = S.queue[head++], r = cur >> 6, c = cur & 63;
        const int dr[4] = {0, 0, -1, 1}, dc[4] = {-1, 1, 0, 0};
        for (int e = 0; e < 4; e++) {
            const int nr = r + dr[e], nc = c + dc[e];
            if (nr < 0 || nr >= 64 || nc < 0 || nc >= 64) continue;
            if (S.visited[nr * 64 + nc]) continue;
            const uint16_t l = gp_li(R, chunk, nr, nc);
            if (l != 0xffff) return l;
            S.visited[nr * 64 + nc] = 1;
            S.queue[tail++] = (uint16_t)(nr * 64 + nc);
        }
    }
    return 0xffff;
}

// N_PortalReachableFromTile (nav.c:4852)
__device__ bool gp_portal_reachable(const dev_route &R, int chunk, const dev_portal &p, int tr, int tc)
{
    for (int r = p.r0; r <= p.r1; r++)
        for (int c = p.c0; c <= p.c1; c++) {
            const uint16_t l = gp_li(R, chunk, r, c);
            if (l == 0xffff) continue;
            for (int r1 = tr - 1; r1 <= tr + 1; r1++)
                for (int c1 = tc - 1; c1 <= tc + 1; c1++) {
                    if (r1 < 0 || r1 >= 64 || c1 < 0 || c1 >= 64) continue;
                    if (l == gp_li(R, chunk, r1, c1)) return true;
                }
        }
    return false;
}

__device__ __forceinline__ int gp_find(const gp_scratch &S, uint64_t key)       // slot of key, or of the first free slot
{
    uint32_t h = (uint32_t)((key * 0x9E3779B97F4A7C15ull) >> 51) & (GP_HT - 1);
    while (S.hkey[h] != 0 && S.hkey[h] != key) h = (h + 1) & (GP_HT - 1);
    return (int)h;
}
// pq push / pop: pqueue.h:165-208 (strict comparisons, hole sift)
__device__ __forceinline__ bool gp_push(gp_scratch &S, int &size, float prio, uint64_t d)
{
    if (size + 2 >= GP_PQ) return false;
    int curr = size + 1, parent = curr / 2;
    while (curr > 1 && S.pprio[parent] > prio) { S.pprio[curr] = S.pprio[parent]; S.pdata[curr] = S.pdata[parent]; curr = parent; parent /= 2; }
    S.pprio[curr] = prio; S.pdata[curr] = d;
    size++;
    return true;
}
__device__ __forceinline__ uint64_t gp_pop(gp_scratch &S, int &size)
{
    const uint64_t out = S.pdata[1];
    const float xp = S.pprio[size]; const uint64_t xd = S.pdata[size];
    size--;
    int root = 1;
    while (true) {
        const int l = root * 2, r = l + 1;
        int target = 0; float tp = xp;
        if (l <= size && S.pprio[l] < tp) { target = l; tp = S.pprio[l]; }
        if (r <= size && S.pprio[r] < tp) { target = r; tp = S.pprio[r]; }
        if (!target) break;
        S.pprio[root] = S.pprio[target]; S.pdata[root] = S.pdata[target];
        root = target;
    }
    S.pprio[root] = xp; S.pdata[root] = xd;
    return out;
}

// req: {start chunk, start tile_r, start tile_c, end chunk, end tile_r, end tile_c, finish chunk, finish portal}
// out (per search, 4 + 3 * max_hops ints): {status (1 found, 0 none, -1 scratch overflow), nhops, cost bits, 0, hops (chunk, portal, liid)...}
__global__ void k_portal_graph_path(dev_route R, const int32_t *__restrict__ req, int n, gp_scratch *scratch, int32_t *out, int max_hops)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    gp_scratch &S = scratch[i];
    const int32_t *q = req + (size_t)i * 8;
    int32_t *o = out + (size_t)i * (4 + 3 * max_hops);
    o[0] = 0; o[1] = 0; o[2] = 0; o[3] = 0;
    const int bchunk = q[0], echunk = q[3], fin_chunk = q[6], fin_pi = q[7];
    const uint16_t start_liid = gp_closest_liid(R, S, bchunk, q[1], q[2]);
    if (start_liid == 0xffff) return;
    const uint16_t end_liid = gp_closest_liid(R, S, echunk, q[4], q[5]);
    if (end_liid == 0xffff) return;
    for (int k = 0; k < GP_HT; k++) S.hkey[k] = 0;
    int size = 0, nkeys = 0;
    bool overflow = false;
    for (int p = 0; p < R.nports[bchunk]; p++) {
        const dev_portal &P = R.ports[bchunk * 64 + p];
        if (!gp_portal_reachable(R, bchunk, P, q[1], q[2])) continue;
        const uint16_t t = R.travel[(size_t)(R.port_base[bchunk] + p) * 4096 + q[1] * 64 + q[2]];
        if (t == 0xffff) continue;
        const float cost = (float)t / 8;
        const uint64_t key = gp_key(bchunk, p, start_liid);
        const int s = gp_find(S, key);
        if (S.hkey[s] == 0) { S.hkey[s] = key; S.hfrom[s] = 0; nkeys++; }
        S.hcost[s] = cost;
        overflow |= !gp_push(S, size, cost, key);
    }
    const uint64_t last = gp_key(fin_chunk, fin_pi, end_liid);
    while (size > 0 && !overflow) {
        const uint64_t cur = gp_pop(S, size);
        if (cur == last) break;
        const int cchunk = (int)((cur >> 8) & 0xffffff), cpi = (int)(cur & 0xff);
        const uint16_t cli = (uint16_t)(cur >> 32);
        const float base = S.hcost[gp_find(S, cur)];
        auto relax = [&](uint64_t nk, float step) {
            const float new_cost = base + step + R.penalty;
            const int s = gp_find(S, nk);
            if (S.hkey[s] == 0) {
                if (nkeys + 1 >= GP_HT * 3 / 4) { overflow = true; return; }
                S.hkey[s] = nk; nkeys++;
            } else if (!(new_cost < S.hcost[s])) return;
            S.hcost[s] = new_cost; S.hfrom[s] = cur;
            overflow |= !gp_push(S, size, new_cost, nk);
        };
        // neighbours_portal_graph (a_star.c:212): the chunk's own portals over active edges whose portal touches the island ...
        const dev_portal &P = R.ports[cchunk * 64 + cpi];
        int nout = 0;
        for (int e = 0; e < P.edge_cnt && nout < 256; e++) {
            const dev_edge E = R.edges[P.edge_off + e];
            if (E.es == 1) continue;
            const dev_portal &NP = R.ports[cchunk * 64 + E.nb];
            bool reach = false;
            for (int r = NP.r0; r <= NP.r1 && !reach; r++)
                for (int c = NP.c0; c <= NP.c1; c++)
                    if (gp_li(R, cchunk, r, c) == cli) { reach = true; break; }
            if (!reach) continue;
            relax(gp_key(cchunk, E.nb, cli), E.cost);
            nout++;
        }
        // ... then the islands of the connected chunk that touch this one across the portal (portal_connected_liids, :151)
        const dev_portal &C = R.ports[P.conn_chunk * 64 + P.conn_idx];
        uint16_t conn_liids[256];
        int nconn = 0;
        bool full = false;
        for (int r1 = P.r0; r1 <= P.r1 && !full; r1++)
            for (int c1 = P.c0; c1 <= P.c1 && !full; c1++) {
                if (gp_li(R, cchunk, r1, c1) != cli) continue;
                const int ar = P.chunk_r * 64 + r1, ac = P.chunk_c * 64 + c1;
                const int cand[4][2] = {{ar - 1, ac}, {ar, ac - 1}, {ar, ac + 1}, {ar + 1, ac}};
                for (int k = 0; k < 4; k++) {
                    const int r2 = cand[k][0] - C.chunk_r * 64, c2 = cand[k][1] - C.chunk_c * 64;
                    if (r2 < C.r0 || r2 > C.r1 || c2 < C.c0 || c2 > C.c1) continue;
                    if (nconn == 256) { full = true; break; }
                    const uint16_t nl = gp_li(R, P.conn_chunk, r2, c2);
                    bool contains = false;
                    for (int j = 0; j < nconn; j++) if (conn_liids[j] == nl) { contains = true; break; }
                    if (!contains && nl != 0xffff) conn_liids[nconn++] = nl;
                }
            }
        for (int j = 0; j < nconn && nout < 256; j++, nout++) relax(gp_key(P.conn_chunk, P.conn_idx, conn_liids[j]), 1.0f);
    }
    if (overflow) { o[0] = -1; return; }
    const int ls = gp_find(S, last);
    if (S.hkey[ls] == 0 || S.hfrom[ls] == 0) return;       // came_from holds no entry for the finish node
    int nh = 0;
    for (uint64_t cur = last; cur != 0; cur = S.hfrom[gp_find(S, cur)]) nh++;
    o[1] = nh; o[2] = __float_as_int(S.hcost[ls]);
    if (nh > max_hops) { o[0] = -1; return; }
    int k = nh;
    for (uint64_t cur = last; cur != 0; cur = S.hfrom[gp_find(S, cur)]) {
        k--;
        o[4 + 3 * k] = (int)((cur >> 8) & 0xffffff); o[4 + 3 * k + 1] = (int)(cur & 0xff); o[4 + 3 * k + 2] = (int)(cur >> 32);
    }
    o[0] = 1;
}

static route_dev_state *route_dev(pfnav_ctx *ctx)
{
    if (!ctx->route_dev_state) ctx->route_dev_state = new route_dev_state();
    return (route_dev_state *)ctx->route_dev_state;
}

}   // namespace

void pfnav_route_dev_forget(pfnav_ctx *ctx)
{
    route_dev_state *D = (route_dev_state *)ctx->route_dev_state;
    if (!D) return;
    if (ctx->device >= 0) {
        cudaSetDevice(ctx->device);
        cudaFree(D->d_ports); cudaFree(D->d_nports); cudaFree(D->d_port_base); cudaFree(D->d_edges); cudaFree(D->d_travel);
        cudaFree(D->d_scratch); cudaFree(D->d_req); cudaFree(D->d_out);
    }
    delete D;
    ctx->route_dev_state = nullptr;
}

// device copies of the routing tables of one layer; the travel-cost index follows the costs (re-sent when the layer was
// rebuilt), the portal table + edge states follow every commit (map_epoch)
static int route_dev_sync(pfnav_ctx *ctx, int layer, pfnav_route_layer &RL, route_dev_state *D)
{
    const int chunks = ctx->chunk_w * ctx->chunk_h;
    if (D->layer == layer && D->epoch == ctx->map_epoch && D->generation == RL.generation) return 0;
    std::vector<dev_portal> ports((size_t)chunks * 64);
    std::vector<int32_t> nports(chunks), base(chunks);
    std::vector<dev_edge> edges;
    size_t rows = 0;
    for (int ch = 0; ch < chunks; ch++) {
        const auto &pp = ctx->portals[layer][ch];
        nports[ch] = (int32_t)pp.size(); base[ch] = (int32_t)rows; rows += pp.size();
        for (size_t i = 0; i < pp.size(); i++) {
            dev_portal d;
            d.r0 = pp[i].r0; d.c0 = pp[i].c0; d.r1 = pp[i].r1; d.c1 = pp[i].c1; d.chunk_r = pp[i].chunk_r; d.chunk_c = pp[i].chunk_c;
            d.conn_chunk = pp[i].conn_chunk; d.conn_idx = pp[i].conn_idx;
            d.edge_off = (int32_t)edges.size(); d.edge_cnt = (int32_t)RL.chunks[ch].edges[i].size();
            for (const auto &e : RL.chunks[ch].edges[i]) edges.push_back({e.nb, e.es, e.cost});
            ports[(size_t)ch * 64 + i] = d;
        }
    }
    const bool retravel = D->layer != layer || D->generation != RL.generation || D->travel_rows != rows;
    if (!D->d_ports) {
        PF_CUDA(cudaMalloc(&D->d_ports, ports.size() * sizeof(dev_portal)));
        PF_CUDA(cudaMalloc(&D->d_nports, (size_t)chunks * 4));
        PF_CUDA(cudaMalloc(&D->d_port_base, (size_t)chunks * 4));
    }
    if (D->nedges < edges.size()) {
        cudaFree(D->d_edges); D->d_edges = nullptr;
        PF_CUDA(cudaMalloc(&D->d_edges, std::max<size_t>(edges.size(), 1) * sizeof(dev_edge)));
        D->nedges = edges.size();
    }
    PF_CUDA(cudaMemcpy(D->d_ports, ports.data(), ports.size() * sizeof(dev_portal), cudaMemcpyHostToDevice));
    PF_CUDA(cudaMemcpy(D->d_nports, nports.data(), (size_t)chunks * 4, cudaMemcpyHostToDevice));
    PF_CUDA(cudaMemcpy(D->d_port_base, base.data(), (size_t)chunks * 4, cudaMemcpyHostToDevice));
    if (!edges.empty()) PF_CUDA(cudaMemcpy(D->d_edges, edges.data(), edges.size() * sizeof(dev_edge), cudaMemcpyHostToDevice));
    if (retravel) {
        cudaFree(D->d_travel); D->d_travel = nullptr;
        PF_CUDA(cudaMalloc(&D->d_travel, std::max<size_t>(rows, 1) * 4096 * 2));
        for (int ch = 0; ch < chunks; ch++)
            if (nports[ch])
                PF_CUDA(cudaMemcpy(D->d_travel + (size_t)base[ch] * 4096, RL.chunks[ch].travel.data(), (size_t)nports[ch] * 4096 * 2,
                                   cudaMemcpyHostToDevice));
        D->travel_rows = rows;
    }
    D->layer = layer; D->epoch = ctx->map_epoch; D->generation = RL.generation;
    return 0;
}

// AStar_PortalGraphPath for n searches. req: 8 ints per search {start chunk index, start tile r, c, end chunk index, end tile
// r, c, finish chunk index, finish portal index}; out: (4 + 3 * max_hops) ints per search {status, nhops, cost (float
// bits), 0, then (chunk, portal, local island) per hop, start to finish}; status 1 = path, 0 = none, -1 = the search
// outgrew its scratch (heap 16 k entries / 6 k nodes) or max_hops. on_device != 0: the kernel; 0: the host planner's own
// routine (what pfnav_route_request_path runs), for comparison.
extern "C" int pfnav_route_graph_paths(pfnav_ctx *ctx, int layer, const int32_t *req, int n, int32_t *out, int max_hops, int on_device)
{
    PF_ARG(ctx && req && out && n >= 0 && max_hops > 0, "args");
    pfnav_route_layer *prl = route_layer(ctx, layer);
    PF_ARG(prl, "pfnav_route_build not called");
    const int cw = ctx->chunk_w, chunks = cw * ctx->chunk_h;
    for (int i = 0; i < n; i++) {
        const int32_t *q = req + (size_t)i * 8;
        PF_ARG(q[0] >= 0 && q[0] < chunks && q[3] >= 0 && q[3] < chunks && q[6] >= 0 && q[6] < chunks, "chunk index");
        PF_ARG(q[1] >= 0 && q[1] < 64 && q[2] >= 0 && q[2] < 64 && q[4] >= 0 && q[4] < 64 && q[5] >= 0 && q[5] < 64, "tile");
        PF_ARG(q[7] >= 0 && q[7] < (int)ctx->portals[layer][q[6]].size(), "finish portal");
    }
    const size_t ostride = 4 + 3 * (size_t)max_hops;
    if (!on_device) {
        Router R{ctx, *prl, layer, cw, ctx->chunk_h};
        for (int ch = 0; ch < chunks; ch++) update_edge_states(ctx, *prl, layer, ch);
        for (int i = 0; i < n; i++) {
            const int32_t *q = req + (size_t)i * 8;
            int32_t *o = out + (size_t)i * ostride;
            memset(o, 0, ostride * 4);
            std::vector<Router::hop> path;
            float cost = 0.0f;
            const tdesc s = {q[0] / cw, q[0] % cw, q[1], q[2]}, e = {q[3] / cw, q[3] % cw, q[4], q[5]};
            if (!R.portal_graph_path(s, e, q[6], q[7], path, &cost)) continue;
            o[1] = (int32_t)path.size(); memcpy(&o[2], &cost, 4);
            if ((int)path.size() > max_hops) { o[0] = -1; continue; }
            for (size_t k = 0; k < path.size(); k++) { o[4 + 3 * k] = path[k].chunk; o[4 + 3 * k + 1] = path[k].pi; o[4 + 3 * k + 2] = path[k].liid; }
            o[0] = 1;
        }
        return PFNAV_OK;
    }
    PF_NEED_DEVICE(ctx);
    PF_CUDA(cudaSetDevice(ctx->device));
    { int rc = pfnav_blockers_flush(ctx); if (rc) return rc; }
    for (int ch = 0; ch < chunks; ch++) update_edge_states(ctx, *prl, layer, ch);      // n_update_all_edge_states (nav.c:1787)
    route_dev_state *D = route_dev(ctx);
    D->epoch = ~0ull;                                                                  // edge states were just refreshed
    if (route_dev_sync(ctx, layer, *prl, D)) return PFNAV_ERR_CUDA;
    const int batch = 4096;                                                            // 364 KB of scratch per search in flight
    if (D->scratch_n < std::min(n, batch)) {
        cudaFree(D->d_scratch); D->d_scratch = nullptr;
        PF_CUDA(cudaMalloc(&D->d_scratch, (size_t)std::min(std::max(n, 1), batch) * sizeof(gp_scratch)));
        D->scratch_n = std::min(std::max(n, 1), batch);
    }
    if (D->req_bytes < (size_t)batch * 32) { cudaFree(D->d_req); PF_CUDA(cudaMalloc(&D->d_req, (size_t)batch * 32)); D->req_bytes = (size_t)batch * 32; }
    if (D->out_bytes < (size_t)batch * ostride * 4) { cudaFree(D->d_out); PF_CUDA(cudaMalloc(&D->d_out, (size_t)batch * ostride * 4)); D->out_bytes = (size_t)batch * ostride * 4; }
    dev_route R;
    R.ports = D->d_ports; R.nports = D->d_nports; R.port_base = D->d_port_base; R.edges = D->d_edges; R.travel = D->d_travel;
    R.liid = ctx->d_liid + (size_t)ctx->W64 * ctx->H64 * layer; R.cw = cw; R.chh = ctx->chunk_h; R.W64 = ctx->W64;
    R.penalty = (float)sqrt(pow(64, 2.0f) + pow(64, 2.0f));                          // portal_node_penalty (a_star.c:298)
    PF_CUDA(cudaDeviceSynchronize());
    for (int b0 = 0; b0 < n; b0 += batch) {
        const int nb = std::min(batch, n - b0);
        PF_CUDA(cudaMemcpy(D->d_req, req + (size_t)b0 * 8, (size_t)nb * 32, cudaMemcpyHostToDevice));
        k_portal_graph_path<<<(nb + 31) / 32, 32>>>(R, (const int32_t *)D->d_req, nb, D->d_scratch, (int32_t *)D->d_out, max_hops);
        ctx->launches++;
        PF_CUDA(cudaGetLastError());
        PF_CUDA(cudaMemcpy(out + (size_t)b0 * ostride, D->d_out, (size_t)nb * ostride * 4, cudaMemcpyDeviceToHost));
    }
    return PFNAV_OK;
}
