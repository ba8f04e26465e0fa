/*
 * oracle/ref_harness.c -- TEST INFRASTRUCTURE ONLY (never part of the product path).
 *
 * Thin C entry points (prefix pfref_) around the UNMODIFIED reference sources under
 * /root/reference/src, compiled where they lie by oracle/Makefile into oracle/_ref/libpfref.so.
 * This file #include's the reference's src/game/movement.c so that its `static` hot-path
 * functions (move_velocity_work, point_seek_vpref, find_neighbours ... movement.c:1524-2023,
 * 2768-2828, 3395-3466) are callable, and supplies the handful of engine services those
 * translation units link against (gamestate getters backed by plain arrays, the M_Nav*
 * one-line wrappers of src/map/map.c:555-1320, SDL atomics, scheduler no-ops).
 *
 * No reference source is copied into this repository: the include below resolves to the
 * read-only checkout at build time (-I/root/reference/src).
 *
 * Used by: tests/ (golden-vector generation + parity), bench.py --impl reference and the
 * cpu_baseline leg.  See oracle/README.md.
 */
#define _GNU_SOURCE
#include <sched.h>
#include <unistd.h>
#include <pthread.h>
/* position.c (compiled as is) owns the real G_Pos_Set, which writes the engine's live position index; the harness
 * keeps the live positions in a plain array instead, so movement.c's calls are routed there */
#define G_Pos_Set pfref_live_pos_set
#include "game/movement.c"
#undef G_Pos_Set

#include "navigation/nav_private.h"
#include "navigation/field.h"
#include "navigation/fieldcache.h"

#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

/* ------------------------------------------------------------------------------------------
 * A minimal `struct map`: the reference only ever reaches it through M_* accessors, all of
 * which are provided below (map.c itself is not compiled: it drags in the renderer).
 * ---------------------------------------------------------------------------------------- */
struct map{
    void   *nav_private;
    vec3_t  pos;
    size_t  width, height;          /* in chunks */
    struct tile **chunk_tiles;      /* [height*width] -> tile[32*32] */
};

#define PFREF_EXPORT __attribute__((visibility("default")))

/* khash instantiations that live in game.c / entity.c in the engine */
__KHASH_IMPL(id,     extern, khint32_t, int,    1, kh_int_hash_func, kh_int_hash_equal)
__KHASH_IMPL(range,  extern, khint32_t, float,  1, kh_int_hash_func, kh_int_hash_equal)
__KHASH_IMPL(entity, extern, khint32_t, char,   0, kh_int_hash_func, kh_int_hash_equal)

unsigned long g_frame_idx = 0;

/* ------------------------------------------------------------------------------------------
 * Engine services with real bodies
 * ---------------------------------------------------------------------------------------- */

void M_GetResolution(const struct map *map, struct map_resolution *out)
{
    /* map.c:946 */
    out->chunk_w = map->width;
    out->chunk_h = map->height;
    out->tile_w = TILES_PER_CHUNK_WIDTH;
    out->tile_h = TILES_PER_CHUNK_HEIGHT;
    out->field_w = TILES_PER_CHUNK_WIDTH * X_COORDS_PER_TILE;
    out->field_h = TILES_PER_CHUNK_HEIGHT * Z_COORDS_PER_TILE;
}

vec3_t M_GetCenterPos(const struct map *map)
{
    /* map.c:956 */
    return (vec3_t){
        map->pos.x - (map->width * TILES_PER_CHUNK_WIDTH * X_COORDS_PER_TILE)/2.0f,
        map->pos.y,
        map->pos.z + (map->height * TILES_PER_CHUNK_HEIGHT * Z_COORDS_PER_TILE)/2.0f,
    };
}

vec3_t M_GetPos(const struct map *map) { return map->pos; }

static struct box pfref_map_box(const struct map *map)
{
    return (struct box){
        map->pos.x, map->pos.z,
        map->width * TILES_PER_CHUNK_WIDTH * X_COORDS_PER_TILE,
        map->height * TILES_PER_CHUNK_HEIGHT * Z_COORDS_PER_TILE,
    };
}

bool M_NavPositionPathable(const struct map *map, enum nav_layer layer, vec2_t xz_pos)
{
    /* map.c:817 */
    if(!C_BoxPointIntersection(xz_pos.x, xz_pos.z, pfref_map_box(map)))
        return false;
    return N_PositionPathable(xz_pos, layer, map->nav_private, map->pos);
}

bool M_NavPositionBlocked(const struct map *map, enum nav_layer layer, vec2_t xz_pos)
{
    /* map.c:831 */
    if(!C_BoxPointIntersection(xz_pos.x, xz_pos.z, pfref_map_box(map)))
        return false;
    return N_PositionBlocked(xz_pos, layer, map->nav_private, map->pos);
}

bool M_NavClosestPathable(const struct map *map, enum nav_layer layer, vec2_t xz_src, vec2_t *out)
{
    /* map.c:868 */
    return N_ClosestPathable(map->nav_private, layer, map->pos, xz_src, out);
}

bool M_NavIsMaximallyClose(const struct map *map, enum nav_layer layer, vec2_t xz_pos,
                           vec2_t xz_dest, float tolerance)
{
    /* map.c:1033 */
    return N_IsMaximallyClose(map->nav_private, layer, map->pos, xz_pos, xz_dest, tolerance);
}

bool M_NavIsAdjacentToImpassable(const struct map *map, enum nav_layer layer, vec2_t xz_pos)
{
    /* map.c:1039 */
    return N_IsAdjacentToImpassable(map->nav_private, layer, map->pos, xz_pos);
}

/* Terrain render height (map.c M_HeightAtPoint): not on the navigation path -- every oracle map
 * is evaluated with y = 0 (entity y is only consumed by the renderer). */
float M_HeightAtPoint(const struct map *map, vec2_t xz) { (void)map; (void)xz; return 0.0f; }

vec2_t M_NavDesiredPointSeekVelocity(const struct map *map, dest_id_t id, vec2_t curr_pos, vec2_t xz_dest)
{
    return N_DesiredPointSeekVelocity(id, curr_pos, xz_dest, map->nav_private, map->pos);
}

bool M_NavHasDestLOS(const struct map *map, dest_id_t id, vec2_t curr_pos, vec2_t xz_dest)
{
    return N_HasDestLOS(id, curr_pos, map->nav_private, map->pos, xz_dest);
}

uint32_t G_FlagsGetFrom(khash_t(id) *table, uint32_t uid)
{
    khiter_t k = kh_get(id, table, uid);
    assert(k != kh_end(table));
    return kh_value(table, k);
}

int G_GetFactionIDFrom(khash_t(id) *table, uint32_t uid)
{
    khiter_t k = kh_get(id, table, uid);
    assert(k != kh_end(table));
    return kh_value(table, k);
}

float G_GetSelectionRadiusFrom(khash_t(range) *table, uint32_t uid)
{
    khiter_t k = kh_get(range, table, uid);
    assert(k != kh_end(table));
    return kh_value(table, k);
}

int Entity_NavLayerWithRadius(uint32_t flags, float radius)
{
    /* entity.c:554 */
    bool water = !!(flags & ENTITY_FLAG_WATER);
    bool air = !!(flags & ENTITY_FLAG_AIR);
    if(radius >= 15.0f)
        return water ? NAV_LAYER_WATER_7X7 : air ? NAV_LAYER_AIR_7X7 : NAV_LAYER_GROUND_7X7;
    else if(radius >= 10.0f)
        return water ? NAV_LAYER_WATER_5X5 : air ? NAV_LAYER_AIR_5X5 : NAV_LAYER_GROUND_5X5;
    else if(radius >= 5.0f)
        return water ? NAV_LAYER_WATER_3X3 : air ? NAV_LAYER_AIR_3X3 : NAV_LAYER_GROUND_3X3;
    else
        return water ? NAV_LAYER_WATER_1X1 : air ? NAV_LAYER_AIR_1X1 : NAV_LAYER_GROUND_1X1;
}

/* Diplomacy: a settable war matrix stands in for the game state (game.c). */
static uint8_t s_pfref_war[MAX_FACTIONS][MAX_FACTIONS];
PFREF_EXPORT void pfref_set_war(int a, int b, int at_war) { s_pfref_war[a][b] = s_pfref_war[b][a] = (uint8_t)at_war; }
bool G_GetDiplomacyState(int fac_id_a, int fac_id_b, enum diplomacy_state *out)
{
    if(fac_id_a == fac_id_b) return false;
    *out = s_pfref_war[fac_id_a][fac_id_b] ? DIPLOMACY_STATE_WAR : DIPLOMACY_STATE_PEACE;
    return true;
}
uint16_t G_GetEnemyFactions(int faction_id)
{
    uint16_t ret = 0;
    for(int i = 0; i < MAX_FACTIONS; i++)
        if(i != faction_id && s_pfref_war[i][faction_id]) ret |= (uint16_t)(1u << i);
    return ret;
}

/* Arrival / formation are inactive in every oracle scenario: the reference
 * then falls through to the plain point-seek branch (movement.c:1516, 1752, 1888). */
struct arrival_state *G_ArrivalGroup_ForLayer(const struct arrival_group *g, enum nav_layer layer) { return NULL; }
bool G_Arrival_NeighbourSettling(const struct arrival_unit_state *us, vec2_t pos, float radius) { return false; }

/* Scheduler / SDL / misc no-ops */
bool     Sched_UsingBigStack(void) { return true; }
void     Sched_TryYield(void) {}
uint32_t Sched_ActiveTID(void) { return 0; }
int  SDL_AtomicSet(SDL_atomic_t *a, int v) { int old = a->value; a->value = v; return old; }
int  SDL_AtomicGet(SDL_atomic_t *a) { return a->value; }
SDL_bool SDL_AtomicCAS(SDL_atomic_t *a, int oldval, int newval)
{ return __sync_bool_compare_and_swap(&a->value, oldval, newval) ? SDL_TRUE : SDL_FALSE; }
int  SDL_GetCPUCount(void) { return (int)sysconf(_SC_NPROCESSORS_ONLN); }
Uint32 SDL_GetTicks(void) { return 0; }
void *mi_malloc(size_t n) { return malloc(n); }
void *mi_calloc(size_t c, size_t n) { return calloc(c, n); }
void *mi_realloc(void *p, size_t n) { return realloc(p, n); }
void  mi_free(void *p) { free(p); }
bool E_Global_Register(enum eventtype event, handler_t handler, void *user, int simmask) { return true; }
bool E_Global_Unregister(enum eventtype event, handler_t handler) { return true; }

void pfref_stub_abort(const char *name)
{
    fprintf(stderr, "[pfref] FATAL: unimplemented engine stub '%s' was called\n", name);
    abort();
}

/* ------------------------------------------------------------------------------------------
 * Exported oracle API
 * ---------------------------------------------------------------------------------------- */

static bool s_inited = false;

PFREF_EXPORT int pfref_init(void)
{
    if(s_inited)
        return 1;
    if(!N_Init())
        return 0;
    s_inited = true;
    return 1;
}

/* pathable: [chunk_h*32][chunk_w*32] map-tile flags (1 = flat pathable tile, 0 = not pathable).
 * Every tile is TILETYPE_FLAT, base_height 0 (SURVEY.md 8a "Tile cost"). */
PFREF_EXPORT void *pfref_map_new(int chunk_w, int chunk_h, const uint8_t *pathable,
                                 float map_x, float map_z)
{
    pfref_init();
    struct map *map = calloc(1, sizeof(struct map));
    map->width = chunk_w;
    map->height = chunk_h;
    map->pos = (vec3_t){map_x, 0.0f, map_z};
    map->chunk_tiles = calloc(chunk_w * chunk_h, sizeof(struct tile*));

    const int TW = TILES_PER_CHUNK_WIDTH, TH = TILES_PER_CHUNK_HEIGHT;
    for(int cr = 0; cr < chunk_h; cr++) {
    for(int cc = 0; cc < chunk_w; cc++) {
        struct tile *tiles = calloc(TW * TH, sizeof(struct tile));
        for(int r = 0; r < TH; r++) {
        for(int c = 0; c < TW; c++) {
            size_t gr = cr * TH + r, gc = cc * TW + c;
            struct tile *t = &tiles[r * TW + c];
            t->pathable = pathable[gr * (chunk_w * TW) + gc] != 0;
            t->type = TILETYPE_FLAT;
            t->base_height = 0;
            t->ramp_height = 0;
        }}
        map->chunk_tiles[cr * chunk_w + cc] = tiles;
    }}
    map->nav_private = N_NewCtxForMapData(chunk_w, chunk_h, TW, TH,
        (const struct tile**)map->chunk_tiles, true);
    if(!map->nav_private) {
        free(map);
        return NULL;
    }
    return map;
}

/* Same, but from explicit tile attributes: tiles = int32[chunk_h*32][chunk_w*32][4] =
 * {pathable, type, base_height, ramp_height} in global row-major order. Exercises
 * n_set_cost_for_tile / n_make_cliff_edges (nav.c:267, 431) for every tile type. */
PFREF_EXPORT void *pfref_map_new_tiles(int chunk_w, int chunk_h, const int32_t *attrs,
                                       float map_x, float map_z)
{
    pfref_init();
    struct map *map = calloc(1, sizeof(struct map));
    map->width = chunk_w;
    map->height = chunk_h;
    map->pos = (vec3_t){map_x, 0.0f, map_z};
    map->chunk_tiles = calloc(chunk_w * chunk_h, sizeof(struct tile*));

    const int TW = TILES_PER_CHUNK_WIDTH, TH = TILES_PER_CHUNK_HEIGHT;
    for(int cr = 0; cr < chunk_h; cr++) {
    for(int cc = 0; cc < chunk_w; cc++) {
        struct tile *tiles = calloc(TW * TH, sizeof(struct tile));
        for(int r = 0; r < TH; r++) {
        for(int c = 0; c < TW; c++) {
            size_t gr = cr * TH + r, gc = cc * TW + c;
            const int32_t *a = attrs + (gr * (chunk_w * TW) + gc) * 4;
            struct tile *t = &tiles[r * TW + c];
            t->pathable = a[0] != 0;
            t->type = (enum tiletype)a[1];
            t->base_height = a[2];
            t->ramp_height = a[3];
        }}
        map->chunk_tiles[cr * chunk_w + cc] = tiles;
    }}
    map->nav_private = N_NewCtxForMapData(chunk_w, chunk_h, TW, TH,
        (const struct tile**)map->chunk_tiles, true);
    if(!map->nav_private) {
        free(map);
        return NULL;
    }
    return map;
}

PFREF_EXPORT void pfref_map_free(void *m)
{
    struct map *map = m;
    if(!map) return;
    N_FC_ClearAll(((struct nav_private*)map->nav_private)->fieldcache);
    N_FreeCtx(map->nav_private);
    for(size_t i = 0; i < map->width * map->height; i++)
        free(map->chunk_tiles[i]);
    free(map->chunk_tiles);
    free(map);
}

static struct nav_private *pfref_priv(void *m) { return ((struct map*)m)->nav_private; }

/* kind: 0 cost_base(u8) 1 blockers(u16) 2 islands(u16) 3 local_islands(u16). out is
 * [chunks][64][64] in chunk-row-major order. */
PFREF_EXPORT void pfref_get_field(void *m, int layer, int kind, void *out)
{
    struct nav_private *priv = pfref_priv(m);
    size_t n = priv->width * priv->height;
    for(size_t i = 0; i < n; i++) {
        const struct nav_chunk *ch = &priv->chunks[layer][i];
        switch(kind) {
        case 0: memcpy((uint8_t*)out  + i * 4096, ch->cost_base, 4096); break;
        case 1: memcpy((uint16_t*)out + i * 4096, ch->blockers, 8192); break;
        case 2: memcpy((uint16_t*)out + i * 4096, ch->islands, 8192); break;
        case 3: memcpy((uint16_t*)out + i * 4096, ch->local_islands, 8192); break;
        default: if(kind >= 16 && kind < 16 + MAX_FACTIONS) memcpy((uint8_t*)out + i * 4096, ch->factions[kind - 16], 4096); break;
        }
    }
}

/* Portal table: 10 ints per portal {chunk_r, chunk_c, idx, ep0.r, ep0.c, ep1.r, ep1.c,
 * conn_chunk_idx, conn_portal_idx, num_neighbours}. Returns the number of portals. */
PFREF_EXPORT int pfref_get_portals(void *m, int layer, int32_t *out, int maxout)
{
    struct nav_private *priv = pfref_priv(m);
    int n = 0;
    for(size_t i = 0; i < priv->width * priv->height; i++) {
        const struct nav_chunk *ch = &priv->chunks[layer][i];
        for(size_t p = 0; p < ch->num_portals; p++) {
            const struct portal *port = &ch->portals[p];
            if(n < maxout) {
                int32_t *o = out + n * 10;
                o[0] = port->chunk.r; o[1] = port->chunk.c; o[2] = (int)p;
                o[3] = port->endpoints[0].r; o[4] = port->endpoints[0].c;
                o[5] = port->endpoints[1].r; o[6] = port->endpoints[1].c;
                o[7] = (int)(port->connected >> PORTAL_REF_PORTAL_BITS);
                o[8] = (int)(port->connected & PORTAL_REF_PORTAL_MASK);
                o[9] = (int)port->num_neighbours;
            }
            n++;
        }
    }
    return n;
}

/* Edge table of one portal: 3 values per edge {neighbour_ref, state, cost(float bits)} */
PFREF_EXPORT int pfref_get_portal_edges(void *m, int layer, int chunk_idx, int portal_idx,
                                        uint32_t *out, int maxout)
{
    struct nav_private *priv = pfref_priv(m);
    const struct portal *port = &priv->chunks[layer][chunk_idx].portals[portal_idx];
    int n = 0;
    for(size_t e = 0; e < port->num_neighbours && n < maxout; e++, n++) {
        out[n*3 + 0] = port->edges[e].neighbour;
        out[n*3 + 1] = port->edges[e].es;
        memcpy(&out[n*3 + 2], &port->edges[e].cost, 4);
    }
    return n;
}

static void pfref_flow_pack(const struct flow_field *ff, uint8_t *out)
{
    for(int r = 0; r < FIELD_RES_R; r++)
    for(int c = 0; c < FIELD_RES_C; c++)
        out[r * FIELD_RES_C + c] = ff->field[r][c].dir_idx;
}

static void pfref_flow_unpack(const uint8_t *in, struct flow_field *ff)
{
    for(int r = 0; r < FIELD_RES_R; r++)
    for(int c = 0; c < FIELD_RES_C; c++)
        ff->field[r][c].dir_idx = in[r * FIELD_RES_C + c];
}

static void pfref_los_pack(const struct LOS_field *lf, uint8_t *out)
{
    for(int r = 0; r < FIELD_RES_R; r++)
    for(int c = 0; c < FIELD_RES_C; c++)
        out[r * FIELD_RES_C + c] = (uint8_t)(lf->field[r][c].visible | (lf->field[r][c].wavefront_blocked << 1));
}

static void pfref_los_unpack(const uint8_t *in, struct LOS_field *lf)
{
    for(int r = 0; r < FIELD_RES_R; r++)
    for(int c = 0; c < FIELD_RES_C; c++) {
        lf->field[r][c].visible = in[r * FIELD_RES_C + c] & 1;
        lf->field[r][c].wavefront_blocked = (in[r * FIELD_RES_C + c] >> 1) & 1;
    }
}

/* N_FlowFieldInit (optional) + N_FlowFieldUpdate(TARGET_TILE); field.c:2020-2083.
 * inout: 4096 bytes, one dir_idx per tile. */
PFREF_EXPORT void pfref_flow_field_tile(void *m, int layer, int chunk_r, int chunk_c,
                                        int tile_r, int tile_c, int faction_id, int init, uint8_t *inout)
{
    struct nav_private *priv = pfref_priv(m);
    struct coord chunk = {chunk_r, chunk_c};
    struct flow_field ff;
    if(init) N_FlowFieldInit(chunk, &ff);
    else { ff.chunk = chunk; pfref_flow_unpack(inout, &ff); }
    struct field_target target = { .type = TARGET_TILE, .tile = (struct coord){tile_r, tile_c} };
    N_FlowFieldUpdate(chunk, priv, faction_id, layer, target, priv->unit_query_ctx, &ff);
    pfref_flow_pack(&ff, inout);
}

/* N_FlowFieldUpdate(TARGET_PORTAL) towards portal `portal_idx` of the chunk. */
PFREF_EXPORT void pfref_flow_field_portal(void *m, int layer, int chunk_r, int chunk_c,
                                          int portal_idx, int port_iid, int next_iid,
                                          int faction_id, int init, uint8_t *inout)
{
    struct nav_private *priv = pfref_priv(m);
    struct coord chunk = {chunk_r, chunk_c};
    const struct nav_chunk *ch = &priv->chunks[layer][chunk_r * priv->width + chunk_c];
    const struct portal *port = &ch->portals[portal_idx];
    struct flow_field ff;
    if(init) N_FlowFieldInit(chunk, &ff);
    else { ff.chunk = chunk; pfref_flow_unpack(inout, &ff); }
    struct field_target target = { .type = TARGET_PORTAL, .pd = (struct portal_desc){
        port, (uint16_t)port_iid, n_portal(priv, layer, port->connected), (uint16_t)next_iid } };
    N_FlowFieldUpdate(chunk, priv, faction_id, layer, target, priv->unit_query_ctx, &ff);
    pfref_flow_pack(&ff, inout);
}

/* Repair chain of N_DesiredPointSeekVelocity (nav.c:3508-3554), applied to one field in place.
 * N_FlowFieldUpdateToNearestPathable (field.c:2247): `start` must be a non-passable tile. */
PFREF_EXPORT void pfref_flow_nearest_pathable(void *m, int layer, int chunk_r, int chunk_c,
                                              int start_r, int start_c, uint8_t *inout)
{
    struct nav_private *priv = pfref_priv(m);
    struct flow_field ff;
    ff.chunk = (struct coord){chunk_r, chunk_c};
    pfref_flow_unpack(inout, &ff);
    N_FlowFieldUpdateToNearestPathable(priv, layer, (struct coord){chunk_r, chunk_c},
        (struct coord){start_r, start_c}, FACTION_ID_NONE, priv->unit_query_ctx, &ff);
    pfref_flow_pack(&ff, inout);
}

/* N_FlowFieldUpdateIslandToNearest (field.c:2307). portal_idx < 0: the field's target is the tile
 * (tile_r, tile_c); else the portal (portal_idx, port_iid, next_iid) of the chunk. */
PFREF_EXPORT void pfref_flow_island_to_nearest(void *m, int layer, int chunk_r, int chunk_c,
                                               int tile_r, int tile_c, int portal_idx, int port_iid,
                                               int next_iid, int local_iid, uint8_t *inout)
{
    struct nav_private *priv = pfref_priv(m);
    const struct nav_chunk *ch = &priv->chunks[layer][chunk_r * priv->width + chunk_c];
    struct flow_field ff;
    ff.chunk = (struct coord){chunk_r, chunk_c};
    pfref_flow_unpack(inout, &ff);
    if(portal_idx < 0) {
        ff.target = (struct field_target){ .type = TARGET_TILE, .tile = (struct coord){tile_r, tile_c} };
    }else{
        const struct portal *port = &ch->portals[portal_idx];
        ff.target = (struct field_target){ .type = TARGET_PORTAL, .pd = (struct portal_desc){
            port, (uint16_t)port_iid, n_portal(priv, layer, port->connected), (uint16_t)next_iid } };
    }
    N_FlowFieldUpdateIslandToNearest((uint16_t)local_iid, priv, layer, FACTION_ID_NONE, priv->unit_query_ctx, &ff);
    pfref_flow_pack(&ff, inout);
}

/* Region fields. Tile coordinates cross this boundary as absolute (r, c) = chunk * 64 + tile.
 * N_CellArrivalFieldCreate (field.c:2445) [+ N_CellArrivalFieldUpdateToNearestPathable (field.c:2603)
 * when start_rc != NULL, the way cell_field_fixup_task calls them, formation.c:3171-3176]. */
static struct tile_desc pfref_td(int ar, int ac) { return (struct tile_desc){ar / 64, ac / 64, ar % 64, ac % 64}; }
static size_t pfref_ws_size(int dim) { return (size_t)dim * dim * 64 + 4096; }

PFREF_EXPORT void pfref_cell_arrival_field(void *m, int dim, int layer, int enemies, const int32_t *target_rc,
                                           const int32_t *center_rc, const int32_t *overlay_rc, int noverlay,
                                           const int32_t *start_rc, uint8_t *out)
{
    struct nav_private *priv = pfref_priv(m);
    size_t ws = pfref_ws_size(dim);
    void *work = malloc(ws);
    struct tile_desc *ov = malloc(sizeof(struct tile_desc) * (noverlay ? noverlay : 1));
    for(int i = 0; i < noverlay; i++) ov[i] = pfref_td(overlay_rc[2*i], overlay_rc[2*i+1]);
    struct nav_cell_overlay overlay = { ov, (size_t)noverlay };
    N_CellArrivalFieldCreate(priv, dim, dim, layer, (uint16_t)enemies, pfref_td(target_rc[0], target_rc[1]),
        pfref_td(center_rc[0], center_rc[1]), out, work, ws, &overlay);
    if(start_rc)
        N_CellArrivalFieldUpdateToNearestPathable(priv, dim, dim, layer, (uint16_t)enemies,
            pfref_td(start_rc[0], start_rc[1]), pfref_td(center_rc[0], center_rc[1]), out, work, ws, &overlay);
    free(ov); free(work);
}

/* N_GroupArrivalFieldCreate (field.c:2525): world-space targets and centre */
PFREF_EXPORT void pfref_group_arrival_field(void *m, int dim, int layer, int enemies, const float *targets_xz,
                                            int ntargets, const float *center_xz, const int32_t *overlay_rc,
                                            int noverlay, uint8_t *out)
{
    struct map *map = m;
    struct nav_private *priv = pfref_priv(m);
    size_t ws = pfref_ws_size(dim);
    void *work = malloc(ws);
    struct tile_desc *ov = malloc(sizeof(struct tile_desc) * (noverlay ? noverlay : 1));
    for(int i = 0; i < noverlay; i++) ov[i] = pfref_td(overlay_rc[2*i], overlay_rc[2*i+1]);
    struct nav_cell_overlay overlay = { ov, (size_t)noverlay };
    N_GroupArrivalFieldCreate(priv, dim, dim, layer, (uint16_t)enemies, map->pos, (const vec2_t*)targets_xz,
        ntargets, (vec2_t){center_xz[0], center_xz[1]}, out, work, ws, &overlay);
    free(ov); free(work);
}

/* N_FlowFieldInit + N_FlowFieldUpdate with a TARGET_ZONE target (field.c:2050 -> field_update_zone, :1810);
 * centre in absolute tile coordinates */
PFREF_EXPORT void pfref_flow_field_zone(void *m, int layer, int chunk_r, int chunk_c, int centre_r, int centre_c,
                                        int radius, uint8_t *out)
{
    struct nav_private *priv = pfref_priv(m);
    struct flow_field ff;
    N_FlowFieldInit((struct coord){chunk_r, chunk_c}, &ff);
    struct field_target target = (struct field_target){ .type = TARGET_ZONE,
        .zone = (struct zone_desc){ pfref_td(centre_r, centre_c), (uint16_t)radius } };
    N_FlowFieldUpdate((struct coord){chunk_r, chunk_c}, priv, 0, layer, target, priv->unit_query_ctx, &ff);
    pfref_flow_pack(&ff, out);
}

#ifndef CLAMP
#define CLAMP(a, min, max) (MIN(MAX((a), (min)), (max)))
#endif
/* N_RequestAsyncGroupArrivalField (nav.c:3921) + N_AwaitAsyncFields inline: the zone fields of every chunk in
 * reach are put into the reference's field cache; then N_DesiredGroupArrivalVelocity (nav.c:3561) per position.
 * out_flags bit0 = returned true, bit1 = at_slot. Returns the number of chunk fields built. */
PFREF_EXPORT int pfref_group_arrival_velocity(void *m, int layer, const float *centre_xz, int radius, int n,
                                              const float *pos_xz, float *out_vel, uint8_t *out_flags)
{
    struct map *map = m;
    struct nav_private *priv = pfref_priv(m);
    struct map_resolution res;
    N_GetResolution(priv, &res);
    vec2_t centre = (vec2_t){centre_xz[0], centre_xz[1]};
    struct tile_desc ct;
    int built = 0;
    if(M_Tile_DescForPoint2D(res, map->pos, centre, &ct)) {
        struct field_target target = (struct field_target){ .type = TARGET_ZONE, .zone = (struct zone_desc){ ct, (uint16_t)radius } };
        int reach = 2 * radius;                                   /* nav.c:3945-3951 */
        int gr = ct.chunk_r * (int)res.tile_h + ct.tile_r, gc = ct.chunk_c * (int)res.tile_w + ct.tile_c;
        int min_cr = CLAMP((gr - reach) / (int)res.tile_h, 0, (int)res.chunk_h - 1);
        int max_cr = CLAMP((gr + reach) / (int)res.tile_h, 0, (int)res.chunk_h - 1);
        int min_cc = CLAMP((gc - reach) / (int)res.tile_w, 0, (int)res.chunk_w - 1);
        int max_cc = CLAMP((gc + reach) / (int)res.tile_w, 0, (int)res.chunk_w - 1);
        for(int cr = min_cr; cr <= max_cr; cr++) {
        for(int cc = min_cc; cc <= max_cc; cc++) {
            struct flow_field ff;
            N_FlowFieldInit((struct coord){cr, cc}, &ff);
            N_FlowFieldUpdate((struct coord){cr, cc}, priv, 0, layer, target, priv->unit_query_ctx, &ff);
            N_FC_PutFlowField(priv->fieldcache, N_FlowFieldID((struct coord){cr, cc}, target, layer), &ff);
            built++;
        }}
    }
    for(int i = 0; i < n; i++) {
        vec2_t vel; bool at_slot = false;
        bool ok = N_DesiredGroupArrivalVelocity((vec2_t){pos_xz[2*i], pos_xz[2*i+1]}, priv, layer, map->pos, centre,
            (uint16_t)radius, &vel, &at_slot);
        out_vel[2*i] = vel.x; out_vel[2*i+1] = vel.z;
        out_flags[i] = (uint8_t)((ok ? 1 : 0) | (at_slot ? 2 : 0));
    }
    return built;
}

/* TARGET_ENTITY / TARGET_ENEMIES fields through a nav_unit_query_ctx assembled from the tables pfref_agents_set
 * fills (N_FlowFieldUpdate field.c:2040-2048). Entities are the uploaded agents (uid = index; none is a building,
 * none is dying); factions come from pfref_agents_set_factions, the war matrix from pfref_set_war. */
bool G_GetDiplomacyStateFrom(enum diplomacy_state (*table)[MAX_FACTIONS], int fac_id_a, int fac_id_b,
                             enum diplomacy_state *out)                        /* game.c:2907 */
{
    if(fac_id_a == fac_id_b)
        return false;
    *out = table[fac_id_a][fac_id_b];
    return true;
}

/* Fog of war is engine state: every entity is visible to the requester here (field_enemy_ent's last test,
 * field.c:980). The transforms / bounding boxes only feed that test, so they are placeholders. */
quat_t Entity_GetRotFrom(khash_t(trans) *table, uint32_t uid) { return (quat_t){0.0f, 0.0f, 0.0f, 1.0f}; }
vec3_t Entity_GetScaleFrom(khash_t(trans) *table, uint32_t uid) { return (vec3_t){1.0f, 1.0f, 1.0f}; }
void   Entity_ModelMatrixFrom(vec3_t pos, quat_t rot, vec3_t scale, mat4x4_t *out) { memset(out, 0, sizeof(*out)); }
void   Entity_CurrentOBBFrom(const struct aabb *aabb, mat4x4_t model, vec3_t scale, struct obb *out) { memset(out, 0, sizeof(*out)); }
bool   G_Fog_ObjVisibleFrom(uint32_t *state, bool enabled, uint16_t fac_mask, const struct obb *obb) { return true; }
static khash_t(aabb) *s_pfref_aabbs;
static int s_nagents;        /* defined with the agent tables below */

static struct nav_unit_query_ctx s_pfref_qctx;
static enum diplomacy_state      s_pfref_diptable[MAX_FACTIONS][MAX_FACTIONS];
static khash_t(id)              *s_pfref_dying;

static struct nav_unit_query_ctx *pfref_query_ctx(void)
{
    struct move_gamestate *gs = &s_move_work.gamestate;
    if(!s_pfref_dying) s_pfref_dying = kh_init(id);
    for(int a = 0; a < MAX_FACTIONS; a++)
        for(int b = 0; b < MAX_FACTIONS; b++)
            s_pfref_diptable[a][b] = s_pfref_war[a][b] ? DIPLOMACY_STATE_WAR : DIPLOMACY_STATE_PEACE;
    memset(&s_pfref_qctx, 0, sizeof(s_pfref_qctx));
    s_pfref_qctx.flags = gs->flags;
    s_pfref_qctx.positions = gs->positions;
    s_pfref_qctx.postree = gs->postree;
    s_pfref_qctx.faction_ids = gs->faction_ids;
    s_pfref_qctx.sel_radiuses = gs->sel_radiuses;
    if(s_pfref_aabbs) kh_destroy(aabb, s_pfref_aabbs);
    s_pfref_aabbs = kh_init(aabb);
    for(size_t i = 0; i < s_nagents; i++) {
        int ret;
        khiter_t k = kh_put(aabb, s_pfref_aabbs, (uint32_t)i, &ret);
        memset(&kh_value(s_pfref_aabbs, k), 0, sizeof(struct aabb));
    }
    s_pfref_qctx.aabbs = s_pfref_aabbs;
    s_pfref_qctx.dying_set = s_pfref_dying;
    s_pfref_qctx.diptable = (void*)s_pfref_diptable;
    return &s_pfref_qctx;
}

PFREF_EXPORT void pfref_agents_set_factions(int n, const int32_t *factions)
{
    struct move_gamestate *gs = &s_move_work.gamestate;
    for(int i = 0; i < n; i++) {
        khiter_t k = kh_get(id, gs->faction_ids, (uint32_t)i);
        if(k != kh_end(gs->faction_ids)) kh_value(gs->faction_ids, k) = factions[i];
    }
}

/* kind 0: TARGET_ENTITY, arg = the target entity's uid; kind 1: TARGET_ENEMIES, arg = the requesting faction */
PFREF_EXPORT void pfref_flow_field_entity(void *m, int layer, int chunk_r, int chunk_c, int kind, int arg, uint8_t *out)
{
    struct map *map = m;
    struct nav_private *priv = pfref_priv(m);
    struct flow_field ff;
    N_FlowFieldInit((struct coord){chunk_r, chunk_c}, &ff);
    struct field_target target;
    memset(&target, 0, sizeof(target));
    if(kind == 0) {
        target.type = TARGET_ENTITY;
        target.ent = (struct entity_desc){ (uint32_t)arg, map->pos };
    }else{
        target.type = TARGET_ENEMIES;
        target.enemies.faction_id = arg;
        target.enemies.map_pos = map->pos;
        target.enemies.chunk = (struct coord){chunk_r, chunk_c};
    }
    N_FlowFieldUpdate((struct coord){chunk_r, chunk_c}, priv, arg, layer, target, pfref_query_ctx(), &ff);
    pfref_flow_pack(&ff, out);
}

/* N_LOSFieldCreate; field.c:2085. prev may be NULL (destination chunk). */
PFREF_EXPORT void pfref_los_field(void *m, int layer, int chunk_r, int chunk_c,
                                  int tgt_chunk_r, int tgt_chunk_c, int tgt_tile_r, int tgt_tile_c,
                                  const uint8_t *prev, int prev_chunk_r, int prev_chunk_c,
                                  uint8_t *out)
{
    struct map *map = m;
    struct nav_private *priv = pfref_priv(m);
    struct tile_desc target = {tgt_chunk_r, tgt_chunk_c, tgt_tile_r, tgt_tile_c};
    /* n_dest_id bit packing, nav.c:839-854 */
    dest_id_t id = (((uint32_t)target.chunk_r & 0x3f) << 26) | (((uint32_t)target.chunk_c & 0x3f) << 20)
                 | (((uint32_t)target.tile_r  & 0x3f) << 14) | (((uint32_t)target.tile_c  & 0x3f) <<  8)
                 | (((uint32_t)layer & 0x0f) << 4) | (uint32_t)FACTION_ID_NONE;
    struct LOS_field lf, prev_lf;
    if(prev) {
        prev_lf.chunk = (struct coord){prev_chunk_r, prev_chunk_c};
        pfref_los_unpack(prev, &prev_lf);
    }
    N_LOSFieldCreate(id, (struct coord){chunk_r, chunk_c}, target, priv, map->pos,
        priv->unit_query_ctx, &lf, prev ? &prev_lf : NULL);
    pfref_los_pack(&lf, out);
}

PFREF_EXPORT void pfref_los_field_faction(void *m, int faction_id, int layer, int chunk_r, int chunk_c,
                                  int tgt_chunk_r, int tgt_chunk_c, int tgt_tile_r, int tgt_tile_c,
                                  const uint8_t *prev, int prev_chunk_r, int prev_chunk_c,
                                  uint8_t *out)
{
    struct map *map = m;
    struct nav_private *priv = pfref_priv(m);
    struct tile_desc target = {tgt_chunk_r, tgt_chunk_c, tgt_tile_r, tgt_tile_c};
    /* n_dest_id bit packing, nav.c:839-854 */
    dest_id_t id = (((uint32_t)target.chunk_r & 0x3f) << 26) | (((uint32_t)target.chunk_c & 0x3f) << 20)
                 | (((uint32_t)target.tile_r  & 0x3f) << 14) | (((uint32_t)target.tile_c  & 0x3f) <<  8)
                 | (((uint32_t)layer & 0x0f) << 4) | ((uint32_t)faction_id & 0x0f);
    struct LOS_field lf, prev_lf;
    if(prev) {
        prev_lf.chunk = (struct coord){prev_chunk_r, prev_chunk_c};
        pfref_los_unpack(prev, &prev_lf);
    }
    N_LOSFieldCreate(id, (struct coord){chunk_r, chunk_c}, target, priv, map->pos,
        priv->unit_query_ctx, &lf, prev ? &prev_lf : NULL);
    pfref_los_pack(&lf, out);
}

PFREF_EXPORT int pfref_request_path(void *m, int layer, float sx, float sz, float dx, float dz,
                                    uint32_t *out_dest_id)
{
    struct map *map = m;
    dest_id_t id = DEST_ID_INVALID;
    bool ok = N_RequestPath(map->nav_private, (vec2_t){sx, sz}, (vec2_t){dx, dz}, map->pos, layer, &id);
    *out_dest_id = id;
    return ok;
}

PFREF_EXPORT int pfref_request_path_attacking(void *m, int layer, int faction_id, float sx, float sz, float dx, float dz,
                                              uint32_t *out_dest_id)
{
    struct map *map = m;
    dest_id_t id = DEST_ID_INVALID;
    bool ok = N_RequestPathAttacking(map->nav_private, (vec2_t){sx, sz}, (vec2_t){dx, dz}, faction_id, map->pos, layer, &id);
    *out_dest_id = id;
    return ok;
}

PFREF_EXPORT uint32_t pfref_dest_id(void *m, int layer, float dx, float dz)
{
    struct map *map = m;
    return N_DestIDForPos(map->nav_private, map->pos, (vec2_t){dx, dz}, layer);
}

/* Field-cache readback: returns 1 if (dest,chunk) has a flow field; writes dirs + ffid */
PFREF_EXPORT int pfref_fc_get_flow(void *m, uint32_t dest_id, int chunk_r, int chunk_c,
                                   uint8_t *out, uint64_t *out_ffid)
{
    struct nav_private *priv = pfref_priv(m);
    ff_id_t ffid;
    if(!N_FC_GetDestFFMapping(priv->fieldcache, dest_id, (struct coord){chunk_r, chunk_c}, &ffid))
        return 0;
    const struct flow_field *ff = N_FC_FlowFieldAt(priv->fieldcache, ffid);
    if(!ff)
        return 0;
    pfref_flow_pack(ff, out);
    if(out_ffid) *out_ffid = ffid;
    return 1;
}

PFREF_EXPORT int pfref_fc_get_los(void *m, uint32_t dest_id, int chunk_r, int chunk_c, uint8_t *out)
{
    struct nav_private *priv = pfref_priv(m);
    if(!N_FC_ContainsLOSField(priv->fieldcache, dest_id, (struct coord){chunk_r, chunk_c}))
        return 0;
    const struct LOS_field *lf = N_FC_LOSFieldAt(priv->fieldcache, dest_id, (struct coord){chunk_r, chunk_c});
    pfref_los_pack(lf, out);
    return 1;
}

PFREF_EXPORT void pfref_fc_clear(void *m) { N_FC_ClearAll(pfref_priv(m)->fieldcache); }

/* compute_desired_velocity + compute_los_state for n positions of one flock
 * (movement.c:4129-4180 -> N_DesiredPointSeekVelocity nav.c:3468, N_HasDestLOS nav.c:4026).
 * vdes is sampled at pos, LOS at los_pos (the reference samples LOS at prev_pos). */
PFREF_EXPORT void pfref_desired_velocity(void *m, uint32_t dest_id, int n, const float *pos_xz,
                                         const float *los_pos_xz, float dest_x, float dest_z,
                                         float *out_vdes, uint8_t *out_los)
{
    struct map *map = m;
    vec2_t dest = {dest_x, dest_z};
    for(int i = 0; i < n; i++) {
        vec2_t p = {pos_xz[2*i], pos_xz[2*i+1]};
        vec2_t lp = {los_pos_xz[2*i], los_pos_xz[2*i+1]};
        if(out_los)
            out_los[i] = N_HasDestLOS(dest_id, lp, map->nav_private, map->pos, dest);
        if(out_vdes) {
            vec2_t v = N_DesiredPointSeekVelocity(dest_id, p, dest, map->nav_private, map->pos);
            out_vdes[2*i] = v.x; out_vdes[2*i+1] = v.z;
        }
    }
}

PFREF_EXPORT void pfref_blockers(void *m, int incref, float x, float z, float radius,
                                 int faction_id, uint32_t flags)
{
    struct map *map = m;
    if(incref) N_BlockersIncref((vec2_t){x, z}, radius, faction_id, flags, map->pos, map->nav_private);
    else       N_BlockersDecref((vec2_t){x, z}, radius, faction_id, flags, map->pos, map->nav_private);
}

/* N_BlockersIncrefOBB / DecrefOBB (nav.c:4685): corners_xz = bottom face corners[0], [1], [5], [4] */
PFREF_EXPORT void pfref_blockers_obb(void *m, int incref, const float *corners_xz, int faction_id, uint32_t flags)
{
    struct map *map = m;
    struct obb obb;
    memset(&obb, 0, sizeof(obb));
    const int idx[4] = {0, 1, 5, 4};
    for(int i = 0; i < 4; i++)
        obb.corners[idx[i]] = (vec3_t){corners_xz[2*i], 0.0f, corners_xz[2*i+1]};
    if(incref) N_BlockersIncrefOBB(map->nav_private, faction_id, flags, map->pos, &obb);
    else       N_BlockersDecrefOBB(map->nav_private, faction_id, flags, map->pos, &obb);
}

PFREF_EXPORT void pfref_update(void *m)
{
    struct map *map = m;
    N_Update(map->nav_private);
    N_ApplyDeferredInvalidations();
}

/* G_ClearPath_NewVelocity (clearpath.c:694). dyn/stat: 5 floats each {px,pz,vx,vz,radius} */
PFREF_EXPORT void pfref_clearpath(const float *self5, const float *vpref2,
                                  const float *dyn, int ndyn, const float *stat, int nstat,
                                  float *out2)
{
    struct cp_ent self = { {self5[0], self5[1]}, {self5[2], self5[3]}, self5[4] };
    vec_cp_ent_t vd, vs;
    vec_cp_ent_init(&vd); vec_cp_ent_init(&vs);
    vec_cp_ent_resize(&vd, MAX_NEIGHBOURS > ndyn ? MAX_NEIGHBOURS : ndyn);
    vec_cp_ent_resize(&vs, MAX_NEIGHBOURS > nstat ? MAX_NEIGHBOURS : nstat);
    for(int i = 0; i < ndyn; i++)
        vec_cp_ent_push(&vd, (struct cp_ent){ {dyn[5*i], dyn[5*i+1]}, {dyn[5*i+2], dyn[5*i+3]}, dyn[5*i+4] });
    for(int i = 0; i < nstat; i++)
        vec_cp_ent_push(&vs, (struct cp_ent){ {stat[5*i], stat[5*i+1]}, {stat[5*i+2], stat[5*i+3]}, stat[5*i+4] });
    vec2_t v = G_ClearPath_NewVelocity(self, 0, (vec2_t){vpref2[0], vpref2[1]}, vd, vs, false);
    out2[0] = v.x; out2[1] = v.z;
    vec_cp_ent_destroy(&vd); vec_cp_ent_destroy(&vs);
}

/* ------------------------------------------------------------------------------------------
 * Agent population: fills the movement module's own statics the way move_copy_gamestate
 * (movement.c:3607) + move_do_tick (movement.c:4333-4408) would.
 * ---------------------------------------------------------------------------------------- */

static int        s_nagents = 0;
static vec3_t    *s_live_pos = NULL;     /* what G_Pos_Set writes on the main thread */
static quat_t    *s_live_rot = NULL;
static bg_ent_t   s_pfref_tree;
static bool       s_tree_valid = false;

static void pfref_agents_clear(void)
{
    struct move_gamestate *gs = &s_move_work.gamestate;
    if(gs->flags)        { kh_destroy(id, gs->flags); gs->flags = NULL; }
    if(gs->positions)    { kh_destroy(pos, gs->positions); gs->positions = NULL; }
    if(gs->sel_radiuses) { kh_destroy(range, gs->sel_radiuses); gs->sel_radiuses = NULL; }
    if(gs->faction_ids)  { kh_destroy(id, gs->faction_ids); gs->faction_ids = NULL; }
    if(s_tree_valid)     { bg_ent_destroy(&s_pfref_tree); s_tree_valid = false; }
    if(s_entity_state_table) { kh_destroy(state, s_entity_state_table); s_entity_state_table = NULL; }
    for(int i = 0; i < vec_size(&s_flocks); i++)
        kh_destroy(entity, vec_AT(&s_flocks, i).ents);
    vec_flock_reset(&s_flocks);
    for(size_t i = 0; i < s_move_work.nwork; i++) {
        vec_cp_ent_destroy(s_move_work.in[i].dyn_neighbs);  free(s_move_work.in[i].dyn_neighbs);
        vec_cp_ent_destroy(s_move_work.in[i].stat_neighbs); free(s_move_work.in[i].stat_neighbs);
    }
    free(s_move_work.in);  s_move_work.in = NULL;
    free(s_move_work.out); s_move_work.out = NULL;
    s_move_work.nwork = 0;
    s_nagents = 0;
}

static bool pfref_uids_equal(const uint32_t *a, const uint32_t *b) { return *a == *b; }

/* uid == agent index. Position index: insert in uid order then cleanup (== G_Pos_Set per
 * entity followed by G_Pos_CopyBitmapGrid, position.c:359), so in-cell order is descending uid.
 * flock_of[i] = flock index (>= 0) ; flock_target: 2 floats per flock ; flock_dest: dest_id */
PFREF_EXPORT void pfref_agents_set(void *m, int n, const float *pos_xz, const float *prev_pos_xz,
                                   const float *vel_xz, const float *radius, const float *max_speed,
                                   const int32_t *state, const uint32_t *flags, const int32_t *flock_of,
                                   int nflocks, const float *flock_target, const uint32_t *flock_dest,
                                   int hz)
{
    struct map *map = m;
    pfref_agents_clear();
    struct move_gamestate *gs = &s_move_work.gamestate;
    gs->flags = kh_init(id);
    gs->positions = kh_init(pos);
    gs->sel_radiuses = kh_init(range);
    gs->faction_ids = kh_init(id);
    gs->map = map;
    s_map = map;
    s_entity_state_table = kh_init(state);
    kh_resize(id, gs->flags, n); kh_resize(pos, gs->positions, n);
    kh_resize(range, gs->sel_radiuses, n); kh_resize(id, gs->faction_ids, n);
    kh_resize(state, s_entity_state_table, n);
    s_move_work.hz = (hz == 20) ? MOVE_HZ_20 : (hz == 10) ? MOVE_HZ_10 : (hz == 5) ? MOVE_HZ_5 : MOVE_HZ_1;

    vec3_t center = M_GetCenterPos(map);
    float hw = (map->width  * TILES_PER_CHUNK_WIDTH  * X_COORDS_PER_TILE) / 2.0f;
    float hh = (map->height * TILES_PER_CHUNK_HEIGHT * Z_COORDS_PER_TILE) / 2.0f;
    bg_ent_init(&s_pfref_tree, center.x - hw, center.x + hw, center.z - hh, center.z + hh, pfref_uids_equal);
    bg_ent_reserve(&s_pfref_tree, n);
    s_tree_valid = true;

    vec_flock_init(&s_flocks);
    for(int f = 0; f < nflocks; f++) {
        struct flock fl;
        memset(&fl, 0, sizeof(fl));
        fl.ents = kh_init(entity);
        fl.target_xz = (vec2_t){flock_target[2*f], flock_target[2*f+1]};
        fl.dest_id = flock_dest[f];
        vec_flock_push(&s_flocks, fl);
    }

    int ret;
    for(int i = 0; i < n; i++) {
        uint32_t uid = i;
        khiter_t k;
        k = kh_put(id, gs->flags, uid, &ret);          kh_value(gs->flags, k) = flags[i];
        k = kh_put(id, gs->faction_ids, uid, &ret);    kh_value(gs->faction_ids, k) = 0;
        k = kh_put(range, gs->sel_radiuses, uid, &ret); kh_value(gs->sel_radiuses, k) = radius[i];
        k = kh_put(pos, gs->positions, uid, &ret);
        kh_value(gs->positions, k) = (vec3_t){pos_xz[2*i], 0.0f, pos_xz[2*i+1]};
        bg_ent_insert(&s_pfref_tree, pos_xz[2*i], pos_xz[2*i+1], uid);

        struct movestate ms;
        memset(&ms, 0, sizeof(ms));
        ms.state = state[i];
        ms.max_speed = max_speed[i];
        ms.velocity = (vec2_t){vel_xz[2*i], vel_xz[2*i+1]};
        ms.prev_pos = (vec3_t){prev_pos_xz[2*i], 0.0f, prev_pos_xz[2*i+1]};
        k = kh_put(state, s_entity_state_table, uid, &ret);
        kh_value(s_entity_state_table, k) = ms;

        if(flock_of[i] >= 0)
            kh_put(entity, vec_AT(&s_flocks, flock_of[i]).ents, uid, &ret);
    }
    bg_ent_cleanup(&s_pfref_tree);
    gs->postree = &s_pfref_tree;
    s_nagents = n;
    free(s_live_pos); free(s_live_rot);
    s_live_pos = malloc(sizeof(vec3_t) * (n ? n : 1));
    s_live_rot = malloc(sizeof(quat_t) * (n ? n : 1));
    for(int i = 0; i < n; i++) {
        s_live_pos[i] = (vec3_t){pos_xz[2*i], 0.0f, pos_xz[2*i+1]};
        s_live_rot[i] = (quat_t){0.0f, 0.0f, 0.0f, 1.0f};
    }
    static bool s_args_inited = false;
    if(!s_args_inited) { stalloc_init(&s_eventargs); s_args_inited = true; }
}

/* ------------------------------------------------------------------------------------------
 * Multi-tick driving of the reference (tests of the device-resident tick loop): the engine services
 * entity_apply_update (movement.c:2693) and entity_finish_moving / entity_block (movement.c:685, 580) call, backed by
 * plain arrays. The "live" position table is what G_Pos_Set writes on the main thread; pfref_publish() is the
 * next tick's move_copy_gamestate (movement.c:3607) + N_Update.
 * ---------------------------------------------------------------------------------------- */
bool G_EntityExists(uint32_t uid) { return (int)uid < s_nagents; }
bool G_EntityIsZombie(uint32_t uid) { (void)uid; return false; }
bool G_EntityIsGarrisoned(uint32_t uid) { return !!(G_FlagsGetFrom(s_move_work.gamestate.flags, uid) & ENTITY_FLAG_GARRISONED); }
uint32_t G_FlagsGet(uint32_t uid) { return G_FlagsGetFrom(s_move_work.gamestate.flags, uid); }
bool pfref_live_pos_set(uint32_t uid, vec3_t pos) { s_live_pos[uid] = pos; return true; }
void Entity_SetRot(uint32_t uid, quat_t rot) { s_live_rot[uid] = rot; }
quat_t Entity_GetRot(uint32_t uid) { return s_live_rot[uid]; }
void E_Entity_Notify(enum eventtype e, uint32_t uid, void *arg, enum event_source src) { (void)e; (void)uid; (void)arg; (void)src; }
void E_Global_Notify(enum eventtype e, void *arg, enum event_source src) { (void)e; (void)arg; (void)src; }
void G_Combat_SetStance(uint32_t uid, enum combat_stance stance) { (void)uid; (void)stance; }
void M_NavBlockersIncref(vec2_t xz_pos, float range, int faction_id, uint32_t flags, const struct map *map)
{ N_BlockersIncref(xz_pos, range, faction_id, flags, map->pos, map->nav_private); }      /* map.c:1043 */
void M_NavBlockersDecref(vec2_t xz_pos, float range, int faction_id, uint32_t flags, const struct map *map)
{ N_BlockersDecref(xz_pos, range, faction_id, flags, map->pos, map->nav_private); }
void M_NavInvalidateZoneFieldsAt(const struct map *map, vec2_t xz_pos, enum nav_layer layer)
{ N_InvalidateZoneFieldsAt(map->nav_private, map->pos, xz_pos, layer); }

/* Query order probe: G_Pos_EntsInCircleFrom (position.c:379) */
PFREF_EXPORT int pfref_ents_in_circle(float x, float z, float range, uint32_t *out, int maxout)
{
    return G_Pos_EntsInCircleFrom(s_move_work.gamestate.postree, s_move_work.gamestate.flags,
        (vec2_t){x, z}, range, out, maxout);
}

/* Build the work list for the given uids the way move_do_tick does (movement.c:4333-4408);
 * vdes/has_los per work item come from pfref_desired_velocity or are supplied directly. */
PFREF_EXPORT void pfref_work_set(int nwork, const uint32_t *uids, const float *vdes,
                                 const uint8_t *has_los, const float *speed)
{
    for(size_t i = 0; i < s_move_work.nwork; i++) {
        vec_cp_ent_destroy(s_move_work.in[i].dyn_neighbs);  free(s_move_work.in[i].dyn_neighbs);
        vec_cp_ent_destroy(s_move_work.in[i].stat_neighbs); free(s_move_work.in[i].stat_neighbs);
    }
    free(s_move_work.in); free(s_move_work.out);
    s_move_work.in = calloc(nwork, sizeof(struct move_work_in));
    s_move_work.out = calloc(nwork, sizeof(struct move_work_out));
    s_move_work.nwork = nwork;
    for(int i = 0; i < nwork; i++) {
        uint32_t uid = uids[i];
        const struct movestate *ms = movestate_get(uid);
        struct move_work_in *in = &s_move_work.in[i];
        in->ent_uid = uid;
        in->ent_des_v = (vec2_t){vdes[2*i], vdes[2*i+1]};
        in->speed = speed[i];
        in->has_dest_los = has_los[i];
        in->fstate.fid = NULL_FID;
        in->cp_ent = (struct cp_ent){
            .xz_pos = (vec2_t){ms->prev_pos.x, ms->prev_pos.z},
            .xz_vel = ms->velocity,
            .radius = G_GetSelectionRadiusFrom(s_move_work.gamestate.sel_radiuses, uid)
        };
        in->dyn_neighbs = malloc(sizeof(vec_cp_ent_t));
        in->stat_neighbs = malloc(sizeof(vec_cp_ent_t));
        vec_cp_ent_init(in->dyn_neighbs);  vec_cp_ent_resize(in->dyn_neighbs, MAX_NEIGHBOURS);
        vec_cp_ent_init(in->stat_neighbs); vec_cp_ent_resize(in->stat_neighbs, MAX_NEIGHBOURS);
    }
}

/* Formation inputs of the work items (struct formation_state + cell_pos + cell_arrival_vdes of struct move_work_in,
 * movement.c:215-225, 264-276): what formation.c's G_Formation_* getters hand to move_push_work (movement.c:4361-4400).
 * form: 14 floats per item {cell_pos[2], cell_arrival_vdes[2], cohesion[2], align[2], drag[2], target_orientation[4]};
 * flags: bit0 has formation, bit1 assignment_ready, bit2 assigned_to_cell, bit3 in_range_of_cell, bit4 arrived_at_cell.
 * For STATE_ARRIVING_TO_CELL the desired velocity of the item is cell_arrival_vdes (ent_desired_velocity, movement.c:1507). */
PFREF_EXPORT void pfref_work_set_formation(int nwork, const float *form, const uint32_t *flags)
{
    for(int i = 0; i < nwork; i++) {
        struct move_work_in *in = &s_move_work.in[i];
        const float *f = form + 14 * i;
        in->cell_pos = (vec2_t){f[0], f[1]};
        in->cell_arrival_vdes = (vec2_t){f[2], f[3]};
        in->fstate.fid = (flags[i] & 1) ? 1 : NULL_FID;
        in->fstate.assignment_ready = !!(flags[i] & 2);
        in->fstate.assigned_to_cell = !!(flags[i] & 4);
        in->fstate.in_range_of_cell = !!(flags[i] & 8);
        in->fstate.arrived_at_cell = !!(flags[i] & 16);
        in->fstate.normal_cohesion_force = (vec2_t){f[4], f[5]};
        in->fstate.normal_align_force = (vec2_t){f[6], f[7]};
        in->fstate.normal_drag_force = (vec2_t){f[8], f[9]};
        in->fstate.target_orientation = (quat_t){f[10], f[11], f[12], f[13]};
        if(movestate_get(in->ent_uid)->state == STATE_ARRIVING_TO_CELL)
            in->ent_des_v = in->cell_arrival_vdes;
    }
}

/* The rest of struct movestate (movement.c:146-215) for the states beyond point seeking, by uid: ints[4] =
 * {wait_prev, wait_ticks_left, surround_target_uid, using_surround_field}, floats[11] = {target_range,
 * target_prev_pos[2], target_dir[4], rot[4]} (rot = what Entity_GetRot returns). */
PFREF_EXPORT void pfref_movestate_ext_set(int n, const int32_t *ints, const float *floats)
{
    for(int i = 0; i < n; i++) {
        struct movestate *ms = movestate_get(i);
        ms->wait_prev = ints[4*i]; ms->wait_ticks_left = ints[4*i+1];
        ms->surround_target_uid = (uint32_t)ints[4*i+2]; ms->using_surround_field = !!ints[4*i+3];
        const float *f = floats + 11 * i;
        ms->target_range = f[0];
        ms->target_prev_pos = (vec2_t){f[1], f[2]};
        ms->target_dir = (quat_t){f[3], f[4], f[5], f[6]};
        s_live_rot[i] = (quat_t){f[7], f[8], f[9], f[10]};
    }
}

/* patch fields beyond the first 25 floats: next_dest[2], next_target_prev[2], next_target_dir[4], next_attack */
PFREF_EXPORT void pfref_compute_updates_ext(int nwork, const float *new_vel, int32_t *out_i, float *out_f, float *out_x)
{
    for(int i = 0; i < nwork; i++) {
        const struct move_work_in *in = &s_move_work.in[i];
        struct movestate_patch p;
        memset(&p, 0, sizeof(p));
        entity_compute_update(s_move_work.hz, in->ent_uid, (vec2_t){new_vel[2*i], new_vel[2*i+1]}, in->ent_des_v, in, &p);
        float *ox = out_x + 9*i;
        memset(ox, 0, 9 * sizeof(float));
        if(p.flags & UPDATE_SET_DEST) { ox[0] = p.next_dest.x; ox[1] = p.next_dest.z; ox[8] = p.next_attack; }
        if(p.flags & UPDATE_SET_TARGET_PREV) { ox[2] = p.next_target_prev.x; ox[3] = p.next_target_prev.z; }
        if(p.flags & UPDATE_SET_TARGET_DIR) { ox[4] = p.next_target_dir.x; ox[5] = p.next_target_dir.y; ox[6] = p.next_target_dir.z; ox[7] = p.next_target_dir.w; }
        /* the first 25 floats + ints exactly as pfref_compute_updates lays them out */
        const struct movestate *ms = movestate_get(in->ent_uid);
        int32_t *oi = out_i + 4*i;
        float *of = out_f + 28*i;
        memset(of, 0, 28 * sizeof(float));
        oi[0] = p.flags;
        oi[1] = (p.flags & (UPDATE_SET_STATE | UPDATE_SET_MOVING)) ? (int32_t)p.next_state : -1;
        oi[2] = (p.flags & UPDATE_SET_STATE) ? (int32_t)p.next_block : 0;
        oi[3] = ms->wait_ticks_left;
        if(p.flags & UPDATE_SET_VELOCITY) { of[0] = p.next_velocity.x; of[1] = p.next_velocity.z; }
        if(p.flags & UPDATE_SET_POSITION) { of[2] = p.next_pos.x; of[3] = p.next_pos.y; of[4] = p.next_pos.z; }
        if(p.flags & UPDATE_SET_ROTATION) { of[5] = p.next_rot.x; of[6] = p.next_rot.y; of[7] = p.next_rot.z; of[8] = p.next_rot.w; }
        if(p.flags & UPDATE_SET_PREV_POS) { of[9] = p.next_ppos.x; of[10] = p.next_ppos.y; of[11] = p.next_ppos.z; }
        if(p.flags & UPDATE_SET_NEXT_POS) { of[12] = p.next_npos.x; of[13] = p.next_npos.y; of[14] = p.next_npos.z; }
        if(p.flags & UPDATE_SET_STEP) of[15] = p.next_step;
        if(p.flags & UPDATE_SET_LEFT) of[16] = p.next_left;
        if(p.flags & UPDATE_SET_NEXT_ROT) { of[17] = p.next_nrot.x; of[18] = p.next_nrot.y; of[19] = p.next_nrot.z; of[20] = p.next_nrot.w; }
        if(p.flags & UPDATE_SET_PREV_ROT) { of[21] = p.next_prot.x; of[22] = p.next_prot.y; of[23] = p.next_prot.z; of[24] = p.next_prot.w; }
    }
}

/* Interpolation / orientation / wait part of `struct movestate` (movement.c:150-215) for the
 * state-update pass. Arrays are indexed by uid; quaternions are {x, y, z, w}. */
PFREF_EXPORT void pfref_movestate_set(int n, const float *next_pos_xz, const float *next_rot,
                                      const float *step, const int32_t *left,
                                      const float *vel_hist, const int32_t *vel_hist_idx,
                                      const int32_t *wait_prev, const int32_t *wait_ticks,
                                      const float *combat_facing)
{
    for(int i = 0; i < n; i++) {
        struct movestate *ms = movestate_get(i);
        ms->next_pos = (vec3_t){next_pos_xz[2*i], 0.0f, next_pos_xz[2*i+1]};
        ms->next_rot = (quat_t){next_rot[4*i], next_rot[4*i+1], next_rot[4*i+2], next_rot[4*i+3]};
        ms->prev_rot = ms->next_rot;
        ms->step = step[i];
        ms->left = left[i];
        for(int k = 0; k < VEL_HIST_LEN; k++)
            ms->vel_hist[k] = (vec2_t){vel_hist[(i*VEL_HIST_LEN + k)*2], vel_hist[(i*VEL_HIST_LEN + k)*2 + 1]};
        ms->vel_hist_idx = vel_hist_idx[i];
        ms->wait_prev = wait_prev[i];
        ms->wait_ticks_left = wait_ticks[i];
        ms->combat_facing = (quat_t){combat_facing[4*i], combat_facing[4*i+1], combat_facing[4*i+2], combat_facing[4*i+3]};
        ms->surround_target_uid = NULL_UID;
    }
}

/* entity_compute_update (movement.c:2303) for every work item, the way move_update_work
 * (movement.c:3468) calls it: new velocity + desired velocity in, `struct movestate_patch` out.
 * Per item: out_i[4] = {flags, next_state, next_block, wait_ticks_left after the call};
 * out_f[28] = next_velocity[2] next_pos[3] next_rot[4] next_ppos[3] next_npos[3] next_step
 *             next_left next_nrot[4] next_prot[4] (3 pad). Fields the patch flags do not select are
 * zeroed so that the comparison is deterministic. */
PFREF_EXPORT void pfref_compute_updates(int nwork, const float *new_vel, int32_t *out_i, float *out_f)
{
    for(int i = 0; i < nwork; i++) {
        const struct move_work_in *in = &s_move_work.in[i];
        struct movestate_patch p;
        memset(&p, 0, sizeof(p));
        entity_compute_update(s_move_work.hz, in->ent_uid, (vec2_t){new_vel[2*i], new_vel[2*i+1]},
                              in->ent_des_v, in, &p);
        const struct movestate *ms = movestate_get(in->ent_uid);
        int32_t *oi = out_i + 4*i;
        float *of = out_f + 28*i;
        memset(of, 0, 28 * sizeof(float));
        oi[0] = p.flags;
        oi[1] = (p.flags & (UPDATE_SET_STATE | UPDATE_SET_MOVING)) ? (int32_t)p.next_state : -1;
        oi[2] = (p.flags & UPDATE_SET_STATE) ? (int32_t)p.next_block : 0;
        oi[3] = ms->wait_ticks_left;
        if(p.flags & UPDATE_SET_VELOCITY) { of[0] = p.next_velocity.x; of[1] = p.next_velocity.z; }
        if(p.flags & UPDATE_SET_POSITION) { of[2] = p.next_pos.x; of[3] = p.next_pos.y; of[4] = p.next_pos.z; }
        if(p.flags & UPDATE_SET_ROTATION) { of[5] = p.next_rot.x; of[6] = p.next_rot.y; of[7] = p.next_rot.z; of[8] = p.next_rot.w; }
        if(p.flags & UPDATE_SET_PREV_POS) { of[9] = p.next_ppos.x; of[10] = p.next_ppos.y; of[11] = p.next_ppos.z; }
        if(p.flags & UPDATE_SET_NEXT_POS) { of[12] = p.next_npos.x; of[13] = p.next_npos.y; of[14] = p.next_npos.z; }
        if(p.flags & UPDATE_SET_STEP) of[15] = p.next_step;
        if(p.flags & UPDATE_SET_LEFT) of[16] = p.next_left;
        if(p.flags & UPDATE_SET_NEXT_ROT) { of[17] = p.next_nrot.x; of[18] = p.next_nrot.y; of[19] = p.next_nrot.z; of[20] = p.next_nrot.w; }
        if(p.flags & UPDATE_SET_PREV_ROT) { of[21] = p.next_prot.x; of[22] = p.next_prot.y; of[23] = p.next_prot.z; of[24] = p.next_prot.w; }
    }
}

/* The movestate part of entity_apply_update (movement.c:2693-2757) for moving entities: velocity +
 * velocity history (update_vel_hist / seed_vel_hist_facing, movement.c:2025-2052), interpolation
 * fields, rotation. State transitions (entity_finish_moving: blockers, events) and G_Pos_Set stay
 * with the engine. Returns the resulting history so that a device-side apply can be checked.
 * out per uid: vel_hist[28], vel_hist_idx */
PFREF_EXPORT void pfref_apply_velocity_patch(int nwork, const float *next_velocity, const int32_t *flags,
                                             float *out_hist, int32_t *out_idx)
{
    for(int i = 0; i < nwork; i++) {
        struct movestate *ms = movestate_get(s_move_work.in[i].ent_uid);
        if(flags[i] & UPDATE_SET_VELOCITY) {
            ms->velocity = (vec2_t){next_velocity[2*i], next_velocity[2*i+1]};
            if(flags[i] & UPDATE_TURNING_IN_PLACE) {
                memset(ms->vel_hist, 0, sizeof(ms->vel_hist));
            }else{
                if(vel_hist_empty(ms) && PFM_Vec2_Len(&ms->velocity) > EPSILON)
                    seed_vel_hist_facing(ms);
                update_vel_hist(ms, ms->velocity);
            }
        }
        for(int k = 0; k < VEL_HIST_LEN; k++) {
            out_hist[(i*VEL_HIST_LEN + k)*2] = ms->vel_hist[k].x;
            out_hist[(i*VEL_HIST_LEN + k)*2 + 1] = ms->vel_hist[k].z;
        }
        out_idx[i] = ms->vel_hist_idx;
    }
}

/* A2 of the navigation tick on the current work list, with the reference's own functions: compute_los_state
 * (movement.c:4129) and compute_desired_velocity (:4163), i.e. N_HasDestLOS + N_DesiredPointSeekVelocity incl. the
 * on-miss chain (inline n_request_path + field repairs). out_vdes / out_los may be NULL. */
PFREF_EXPORT void pfref_desired_from_cache(float *out_vdes, uint8_t *out_los)
{
    compute_los_state();
    compute_desired_velocity();
    for(size_t i = 0; i < s_move_work.nwork; i++) {
        if(out_vdes) { out_vdes[2*i] = s_move_work.in[i].ent_des_v.x; out_vdes[2*i+1] = s_move_work.in[i].ent_des_v.z; }
        if(out_los) out_los[i] = s_move_work.in[i].has_dest_los;
    }
}

/* entity_compute_update + entity_apply_update (movement.c:2303, 2693) for every work item with the velocities the
 * velocity pass left in the work outputs: the real functions, engine services above. Then the next tick's
 * snapshot: positions table + position index rebuilt from the live table, N_Update for blockers taken by
 * entity_block. Returns the number of entities whose state changed. */
PFREF_EXPORT int pfref_update_and_apply(void *m)
{
    int nwork = (int)s_move_work.nwork, changed = 0;
    struct movestate_patch *patches = calloc(nwork ? nwork : 1, sizeof(struct movestate_patch));
    for(int i = 0; i < nwork; i++) {
        const struct move_work_in *in = &s_move_work.in[i];
        entity_compute_update(s_move_work.hz, in->ent_uid, s_move_work.out[i].ent_vel, in->ent_des_v, in, &patches[i]);
    }
    for(int i = 0; i < nwork; i++) {
        uint32_t uid = s_move_work.in[i].ent_uid;
        enum move_state before = movestate_get(uid)->state;
        entity_apply_update(uid, &patches[i]);
        changed += (movestate_get(uid)->state != before);
    }
    free(patches);
    stalloc_clear(&s_eventargs);
    /* publish: move_copy_gamestate + G_Pos_CopyBitmapGrid */
    struct move_gamestate *gs = &s_move_work.gamestate;
    bg_ent_destroy(&s_pfref_tree);
    struct map *map = m;
    vec3_t center = M_GetCenterPos(map);
    float hw = (map->width  * TILES_PER_CHUNK_WIDTH  * X_COORDS_PER_TILE) / 2.0f;
    float hh = (map->height * TILES_PER_CHUNK_HEIGHT * Z_COORDS_PER_TILE) / 2.0f;
    bg_ent_init(&s_pfref_tree, center.x - hw, center.x + hw, center.z - hh, center.z + hh, pfref_uids_equal);
    bg_ent_reserve(&s_pfref_tree, s_nagents);
    for(int i = 0; i < s_nagents; i++) {
        khiter_t k = kh_get(pos, gs->positions, (uint32_t)i);
        kh_value(gs->positions, k) = s_live_pos[i];
        bg_ent_insert(&s_pfref_tree, s_live_pos[i].x, s_live_pos[i].z, (uint32_t)i);
    }
    bg_ent_cleanup(&s_pfref_tree);
    gs->postree = &s_pfref_tree;
    N_Update(map->nav_private);
    N_ApplyDeferredInvalidations();
    return changed;
}

/* entity state after the ticks so far, by uid: pos, prev_pos, velocity (2 floats each), state, blocking */
PFREF_EXPORT void pfref_state_get(int n, float *pos, float *prev_pos, float *vel, int32_t *state, int32_t *blocking)
{
    for(int i = 0; i < n; i++) {
        const struct movestate *ms = movestate_get(i);
        pos[2*i] = s_live_pos[i].x; pos[2*i+1] = s_live_pos[i].z;
        prev_pos[2*i] = ms->prev_pos.x; prev_pos[2*i+1] = ms->prev_pos.z;
        vel[2*i] = ms->velocity.x; vel[2*i+1] = ms->velocity.z;
        state[i] = ms->state; blocking[i] = ms->blocking;
    }
}

/* move_velocity_work over [begin, end] (inclusive), re-entrant over disjoint ranges. */
PFREF_EXPORT void pfref_velocity_work(int begin, int end)
{
    for(int i = begin; i <= end; i++) {
        vec_cp_ent_reset(s_move_work.in[i].dyn_neighbs);
        vec_cp_ent_reset(s_move_work.in[i].stat_neighbs);
    }
    move_velocity_work(begin, end);
}

struct pfref_range { int begin, end; };
static void *pfref_vel_thread(void *arg)
{
    struct pfref_range *r = arg;
    pfref_velocity_work(r->begin, r->end);
    return NULL;
}

/* The reference's static equal-range fork-join (movement.c:3746-3774) on pthreads.
 * Returns elapsed seconds. */
PFREF_EXPORT double pfref_velocity_work_mt(int nthreads)
{
    int nwork = (int)s_move_work.nwork;
    if(nwork == 0) return 0.0;
    if(nthreads > MAX_MOVE_TASKS) nthreads = MAX_MOVE_TASKS;
    if(nwork < 64 || nthreads < 1) nthreads = 1;
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    pthread_t th[MAX_MOVE_TASKS];
    struct pfref_range rg[MAX_MOVE_TASKS];
    size_t nitems = ceil((float)nwork / nthreads);
    int nt = 0;
    for(int i = 0; i < nthreads; i++) {
        int b = nitems * i, e = MIN(nitems * (i + 1) - 1, (size_t)nwork - 1);
        if(b > e) break;
        rg[nt] = (struct pfref_range){b, e};
        pthread_create(&th[nt], NULL, pfref_vel_thread, &rg[nt]);
        nt++;
    }
    for(int i = 0; i < nt; i++) pthread_join(th[i], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    return (t1.tv_sec - t0.tv_sec) + (t1.tv_nsec - t0.tv_nsec) * 1e-9;
}

PFREF_EXPORT void pfref_work_get(int nwork, float *out_vel)
{
    for(int i = 0; i < nwork; i++) {
        out_vel[2*i] = s_move_work.out[i].ent_vel.x;
        out_vel[2*i+1] = s_move_work.out[i].ent_vel.z;
    }
}

/* vpref only (point_seek_vpref, movement.c:1870) for diagnostics */
PFREF_EXPORT void pfref_vpref(int nwork, float *out_vpref)
{
    for(int i = 0; i < nwork; i++) {
        struct move_work_in *in = &s_move_work.in[i];
        const struct flock *flock = flock_for_ent(in->ent_uid);
        vec2_t v = point_seek_vpref(in->ent_uid, flock, in->ent_des_v, in->has_dest_los, in->speed);
        out_vpref[2*i] = v.x; out_vpref[2*i+1] = v.z;
    }
}

/* Timed field batches for the CPU baseline: reqs = 4 ints {chunk_r, chunk_c, tile_r, tile_c}.
 * what: 0 = N_FlowFieldInit+Update(TARGET_TILE), 1 = N_LOSFieldCreate (destination chunk). */
struct pfref_field_job { void *m; int layer; int what; const int32_t *reqs; int begin, end; uint8_t *out; };
static void *pfref_field_thread(void *arg)
{
    struct pfref_field_job *j = arg;
    for(int i = j->begin; i < j->end; i++) {
        const int32_t *q = j->reqs + 4*i;
        uint8_t *o = j->out ? j->out + (size_t)i * 4096 : NULL;
        uint8_t tmp[4096];
        if(j->what == 0)
            pfref_flow_field_tile(j->m, j->layer, q[0], q[1], q[2], q[3], FACTION_ID_NONE, 1, o ? o : tmp);
        else
            pfref_los_field(j->m, j->layer, q[0], q[1], q[0], q[1], q[2], q[3], NULL, 0, 0, o ? o : tmp);
    }
    return NULL;
}

PFREF_EXPORT double pfref_fields_mt(void *m, int layer, int what, const int32_t *reqs, int n,
                                    int nthreads, uint8_t *out)
{
    if(nthreads < 1) nthreads = 1;
    if(nthreads > 256) nthreads = 256;
    struct nav_private *priv = pfref_priv(m);
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    pthread_t th[256];
    struct pfref_field_job jobs[256];
    int per = (n + nthreads - 1) / nthreads, nt = 0;
    for(int i = 0; i < nthreads; i++) {
        int b = per * i, e = MIN(per * (i + 1), n);
        if(b >= e) break;
        jobs[nt] = (struct pfref_field_job){m, layer, what, reqs, b, e, out};
        pthread_attr_t attr;
        pthread_attr_init(&attr);
        pthread_attr_setstacksize(&attr, 8u << 20);   /* TASK_BIG_STACK, sched.c:186 */
        pthread_create(&th[nt], &attr, pfref_field_thread, &jobs[nt]);
        pthread_attr_destroy(&attr);
        nt++;
    }
    for(int i = 0; i < nt; i++) pthread_join(th[i], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    return (t1.tv_sec - t0.tv_sec) + (t1.tv_nsec - t0.tv_nsec) * 1e-9;
}
