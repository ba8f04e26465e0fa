/*
 * shim/field_pfnav.c -- seam B2 of SURVEY.md 8b as a COMPILED drop-in: this translation unit replaces the reference's
 * src/navigation/field.c at link time. It exports every function field.c exports, with the reference's own signatures
 * (src/navigation/field.h:115-202, src/navigation/public/nav.h:700-736 for the arrival fields and N_FlowDir), and
 * implements them on libpfnav.so (include/pfnav.h): the fields are built by the sm_100a kernels.
 *
 * Build: against the reference's headers WHERE THEY LIE (-iquote <reference>/src), never copied; see oracle/Makefile
 * target `shimref`, which links the reference's own nav.c + a_star.c + fieldcache.c (everything except field.c) with
 * this file into oracle/_ref/libpfref_shim.so. tests/test_gpu_shim.py then drives the reference's n_request_path /
 * N_DesiredPointSeekVelocity / N_HasDestLOS through both libraries and requires identical results.
 *
 * State: one device context per `struct nav_private` the engine hands in. The engine owns the navigation state (the
 * caller's `priv` is a read-only snapshot valid for the call), so every call first mirrors the chunks it
 * reads from `priv` into the device images (cost_base, blockers, local_islands of the chunk; the neighbour chunk's
 * islands for TARGET_PORTAL) -- 20 KB per chunk -- then runs the batch entry point with n = 1. An engine that wants
 * throughput calls the batch API directly; this file is the zero-change path.
 * Thread contract: N_FlowFieldUpdate is re-entrant in the reference (parallel field tasks); here a mutex serialises
 * the device calls.
 */
#include "navigation/nav_private.h"
#include "navigation/field.h"
#include "navigation/public/nav.h"

#include "../include/pfnav.h"

#include <assert.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define SHIM_MAX_CTX 8

struct shim_ctx{
    const void *priv_key;           /* identity of the engine context: its chunk buffer */
    size_t      width, height;
    pfnav_ctx  *nav;
    bool        layer_built[NAV_LAYER_MAX];
    float       map_x, map_z;
};

static struct shim_ctx  s_ctxs[SHIM_MAX_CTX];
static pthread_mutex_t  s_lock = PTHREAD_MUTEX_INITIALIZER;

static void shim_die(const char *what)
{
    fprintf(stderr, "[pfnav shim] %s: %s\n", what, pfnav_last_error());
    abort();                        /* the reference's field.c has no error channel either: its failures are asserts */
}

#define SHIM_CHK(call) do { if((call) != PFNAV_OK) shim_die(#call); } while(0)

static const struct nav_chunk *shim_chunk(const struct nav_private *priv, enum nav_layer layer, struct coord c)
{
    return &priv->chunks[layer][c.r * priv->width + c.c];
}

/* the device context that mirrors `priv`; (re)built when the engine hands in a context of another map */
static struct shim_ctx *shim_for(const struct nav_private *priv)
{
    struct shim_ctx *free_slot = NULL;
    for(int i = 0; i < SHIM_MAX_CTX; i++) {
        if(s_ctxs[i].nav && s_ctxs[i].priv_key == (const void*)priv->chunks[0]
        && s_ctxs[i].width == priv->width && s_ctxs[i].height == priv->height)
            return &s_ctxs[i];
        if(!s_ctxs[i].nav && !free_slot)
            free_slot = &s_ctxs[i];
    }
    if(!free_slot) {                /* recycle the oldest */
        pfnav_destroy(s_ctxs[0].nav);
        memmove(&s_ctxs[0], &s_ctxs[1], sizeof(s_ctxs[0]) * (SHIM_MAX_CTX - 1));
        free_slot = &s_ctxs[SHIM_MAX_CTX - 1];
        memset(free_slot, 0, sizeof(*free_slot));
    }
    struct shim_ctx *s = free_slot;
    memset(s, 0, sizeof(*s));
    const char *dev = getenv("PFNAV_DEVICE");
    SHIM_CHK(pfnav_create(dev ? atoi(dev) : 0, &s->nav));
    s->priv_key = priv->chunks[0];
    s->width = priv->width;
    s->height = priv->height;
    SHIM_CHK(pfnav_map_create(s->nav, (int)priv->width, (int)priv->height, NAV_LAYER_MAX, 0.0f, 0.0f));
    return s;
}

/* first use of a layer: the whole layer + the structures derived once per map (islands, portals, routing tables) */
static void shim_build_layer(struct shim_ctx *s, const struct nav_private *priv, enum nav_layer layer)
{
    if(s->layer_built[layer])
        return;
    const size_t nchunks = priv->width * priv->height;
    uint8_t  *cost = malloc(nchunks * 4096);
    uint16_t *blk  = malloc(nchunks * 8192);
    uint16_t *liid = malloc(nchunks * 8192);
    for(size_t i = 0; i < nchunks; i++) {
        const struct nav_chunk *ch = &priv->chunks[layer][i];
        memcpy(cost + i * 4096, ch->cost_base, 4096);
        memcpy(blk + i * 4096, ch->blockers, 8192);
        memcpy(liid + i * 4096, ch->local_islands, 8192);
    }
    SHIM_CHK(pfnav_map_upload_layer(s->nav, layer, cost, blk, liid));
    SHIM_CHK(pfnav_map_build_nav(s->nav, layer));
    SHIM_CHK(pfnav_route_build(s->nav, layer));
    free(cost); free(blk); free(liid);
    s->layer_built[layer] = true;
}

static void shim_sync_chunk(struct shim_ctx *s, const struct nav_private *priv, enum nav_layer layer, struct coord c)
{
    const struct nav_chunk *ch = shim_chunk(priv, layer, c);
    SHIM_CHK(pfnav_map_update_chunk(s->nav, layer, c.r, c.c, &ch->cost_base[0][0], &ch->blockers[0][0],
                                    &ch->local_islands[0][0]));
}

static void shim_sync_factions(struct shim_ctx *s, const struct nav_private *priv, enum nav_layer layer, int faction_id)
{
    if(faction_id == FACTION_ID_NONE)
        return;
    const size_t nchunks = priv->width * priv->height;
    uint8_t *fac = malloc(nchunks * MAX_FACTIONS * 4096);
    for(size_t i = 0; i < nchunks; i++)
        memcpy(fac + i * MAX_FACTIONS * 4096, priv->chunks[layer][i].factions, MAX_FACTIONS * 4096);
    SHIM_CHK(pfnav_map_upload_factions(s->nav, layer, fac));
    free(fac);
    for(int f = 0; f < MAX_FACTIONS; f++)
        SHIM_CHK(pfnav_set_enemy_factions(s->nav, f, G_GetEnemyFactions(f)));
}

/* struct field_target (TILE / PORTAL) -> the request record of the batch API */
static pfnav_field_req shim_req(struct coord chunk, enum nav_layer layer, int faction_id, struct field_target target, int init)
{
    pfnav_field_req q;
    memset(&q, 0, sizeof(q));
    q.chunk_r = chunk.r; q.chunk_c = chunk.c; q.layer = layer; q.faction_id = faction_id; q.init = init;
    if(target.type == TARGET_TILE) {
        q.target_type = PFNAV_TARGET_TILE;
        q.tile_r = target.tile.r; q.tile_c = target.tile.c;
    }else{
        assert(target.type == TARGET_PORTAL);
        const struct portal *p = target.pd.port, *n = target.pd.next;
        q.target_type = PFNAV_TARGET_PORTAL;
        q.port_r0 = p->endpoints[0].r; q.port_c0 = p->endpoints[0].c; q.port_r1 = p->endpoints[1].r; q.port_c1 = p->endpoints[1].c;
        q.next_r0 = n->endpoints[0].r; q.next_c0 = n->endpoints[0].c; q.next_r1 = n->endpoints[1].r; q.next_c1 = n->endpoints[1].c;
        q.next_chunk_r = n->chunk.r; q.next_chunk_c = n->chunk.c;
        q.port_iid = target.pd.port_iid; q.next_iid = target.pd.next_iid;
    }
    return q;
}

static void shim_unpack_flow(const struct flow_field *ff, uint8_t *out)
{
    for(int r = 0; r < FIELD_RES_R; r++)
    for(int c = 0; c < FIELD_RES_C; c++)
        out[r * FIELD_RES_C + c] = ff->field[r][c].dir_idx;
}

static void shim_pack_flow(const uint8_t *in, struct flow_field *ff)
{
    for(int r = 0; r < FIELD_RES_R; r++)
    for(int c = 0; c < FIELD_RES_C; c++)
        ff->field[r][c].dir_idx = in[r * FIELD_RES_C + c] & 0xf;
}

/*****************************************************************************/
/* field.h                                                                   */
/*****************************************************************************/

ff_id_t N_FlowFieldID(struct coord chunk, struct field_target target, enum nav_layer layer)
{
    /* bit layout of field.c:1952-2008 (the ids are keys of the engine's field cache) */
    const uint64_t head = ((uint64_t)layer << 60) | ((uint64_t)target.type << 56);
    const uint64_t tail = ((uint64_t)chunk.r << 8) | (uint64_t)chunk.c;
    switch(target.type) {
    case TARGET_PORTAL:
        return head | tail
             | (((uint64_t)target.pd.next_iid & 0xf) << 48) | (((uint64_t)target.pd.port_iid & 0xf) << 40)
             | ((uint64_t)target.pd.port->endpoints[0].r << 34) | ((uint64_t)target.pd.port->endpoints[0].c << 28)
             | ((uint64_t)target.pd.port->endpoints[1].r << 22) | ((uint64_t)target.pd.port->endpoints[1].c << 16);
    case TARGET_TILE:
        return head | tail | ((uint64_t)target.tile.r << 24) | ((uint64_t)target.tile.c << 16);
    case TARGET_ENEMIES:
        return head | tail | ((uint64_t)target.enemies.faction_id << 24);
    case TARGET_ENTITY:
        return head | tail | ((uint64_t)target.ent.target << 24);
    case TARGET_ZONE:
        return head | tail
             | ((uint64_t)(target.zone.radius & 0xff) << 44)
             | ((uint64_t)(target.zone.centre.tile_c & 0x3f) << 38) | ((uint64_t)(target.zone.centre.tile_r & 0x3f) << 32)
             | ((uint64_t)(target.zone.centre.chunk_c & 0xff) << 24) | ((uint64_t)(target.zone.centre.chunk_r & 0xff) << 16);
    default:
        assert(0);
        return 0;
    }
}

enum nav_layer N_FlowFieldLayer(ff_id_t id) { return (enum nav_layer)(id >> 60); }
int N_FlowFieldTargetType(ff_id_t id) { return (int)((id >> 56) & 0xf); }

void N_FlowFieldInit(struct coord chunk_coord, struct flow_field *out)
{
    memset(out->field, 0, sizeof(out->field));          /* FD_NONE == 0 */
    out->chunk = chunk_coord;
}

void N_FlowFieldUpdate(struct coord chunk_coord, const struct nav_private *priv, int faction_id, enum nav_layer layer,
                       struct field_target target, struct nav_unit_query_ctx *ctx, struct flow_field *inout_flow)
{
    (void)ctx;
    uint8_t bytes[FIELD_RES_R * FIELD_RES_C];
    pthread_mutex_lock(&s_lock);
    struct shim_ctx *s = shim_for(priv);
    shim_build_layer(s, priv, layer);
    shim_sync_chunk(s, priv, layer, chunk_coord);
    if(target.type == TARGET_ZONE) {
        /* field_update_zone (field.c:1810): the integration runs over the chunk padded by half a chunk */
        for(int dr = -1; dr <= 1; dr++)
        for(int dc = -1; dc <= 1; dc++) {
            struct coord nb = {chunk_coord.r + dr, chunk_coord.c + dc};
            if(nb.r < 0 || nb.c < 0 || nb.r >= (int)priv->height || nb.c >= (int)priv->width) continue;
            shim_sync_chunk(s, priv, layer, nb);
        }
        const int32_t chunk_rc[2] = {chunk_coord.r, chunk_coord.c};
        const struct tile_desc ct = target.zone.centre;
        SHIM_CHK(pfnav_zone_fields(s->nav, layer, ct.chunk_r * FIELD_RES_R + ct.tile_r, ct.chunk_c * FIELD_RES_C + ct.tile_c,
                                   target.zone.radius, chunk_rc, 1, bytes));
        inout_flow->target = target;
        shim_pack_flow(bytes, inout_flow);
        pthread_mutex_unlock(&s_lock);
        return;
    }
    if(target.type != TARGET_TILE && target.type != TARGET_PORTAL) {
        /* TARGET_ENEMIES / TARGET_ENTITY read the engine's entity tables through `ctx`; their device entry point takes
         * the footprints directly (pfnav_entity_fields) */
        fprintf(stderr, "[pfnav shim] N_FlowFieldUpdate: target type %d goes through pfnav_entity_fields, not through this shim\n",
                (int)target.type);
        abort();
    }
    if(target.type == TARGET_PORTAL)
        shim_sync_chunk(s, priv, layer, target.pd.next->chunk);
    shim_sync_factions(s, priv, layer, faction_id);
    /* the caller's field is updated IN PLACE (nav.c:1998-2008 merges several targets into one field) */
    pfnav_field_req q = shim_req(chunk_coord, layer, faction_id, target, 0);
    shim_unpack_flow(inout_flow, bytes);
    SHIM_CHK(pfnav_flow_fields_update(s->nav, &q, 1, bytes));
    inout_flow->target = target;                        /* field.c:2076 */
    shim_pack_flow(bytes, inout_flow);
    pthread_mutex_unlock(&s_lock);
}

static void shim_repair(const struct nav_private *priv, enum nav_layer layer, int faction_id, int kind, int32_t arg,
                        struct flow_field *inout_flow)
{
    uint8_t bytes[FIELD_RES_R * FIELD_RES_C];
    pthread_mutex_lock(&s_lock);
    struct shim_ctx *s = shim_for(priv);
    shim_build_layer(s, priv, layer);
    shim_sync_chunk(s, priv, layer, inout_flow->chunk);
    if(inout_flow->target.type == TARGET_PORTAL)
        shim_sync_chunk(s, priv, layer, inout_flow->target.pd.next->chunk);
    shim_sync_factions(s, priv, layer, faction_id);
    pfnav_field_req q = shim_req(inout_flow->chunk, layer, faction_id, inout_flow->target, 0);
    const int32_t k = kind;
    shim_unpack_flow(inout_flow, bytes);
    SHIM_CHK(pfnav_flow_fields_repair(s->nav, &q, &k, &arg, 1, bytes));
    shim_pack_flow(bytes, inout_flow);
    pthread_mutex_unlock(&s_lock);
}

void N_FlowFieldUpdateIslandToNearest(uint16_t local_iid, const struct nav_private *priv, enum nav_layer layer,
                                      int faction_id, struct nav_unit_query_ctx *ctx, struct flow_field *inout_flow)
{
    (void)ctx;
    shim_repair(priv, layer, faction_id, PFNAV_REPAIR_ISLAND_TO_NEAREST, local_iid, inout_flow);
}

void N_FlowFieldUpdateToNearestPathable(const struct nav_private *priv, enum nav_layer layer, struct coord chunk,
                                        struct coord start, int faction_id, struct nav_unit_query_ctx *ctx,
                                        struct flow_field *inout_flow)
{
    (void)ctx; (void)chunk;
    shim_repair(priv, layer, faction_id, PFNAV_REPAIR_NEAREST_PATHABLE, (start.r << 8) | start.c, inout_flow);
}

void N_LOSFieldCreate(dest_id_t id, struct coord chunk_coord, struct tile_desc target, const struct nav_private *priv,
                      vec3_t map_pos, struct nav_unit_query_ctx *ctx, struct LOS_field *out_los,
                      const struct LOS_field *prev_los)
{
    (void)ctx;
    uint8_t bytes[FIELD_RES_R * FIELD_RES_C];
    const enum nav_layer layer = N_DestLayer(id);
    const int faction_id = N_DestFactionID(id);
    pthread_mutex_lock(&s_lock);
    struct shim_ctx *s = shim_for(priv);
    shim_build_layer(s, priv, layer);
    if(s->map_x != map_pos.x || s->map_z != map_pos.z) {
        SHIM_CHK(pfnav_map_set_pos(s->nav, map_pos.x, map_pos.z));      /* the blocked lines are cast in world coordinates */
        s->map_x = map_pos.x; s->map_z = map_pos.z;
    }
    shim_sync_chunk(s, priv, layer, chunk_coord);
    shim_sync_factions(s, priv, layer, faction_id);
    pfnav_los_req q;
    memset(&q, 0, sizeof(q));
    q.chunk_r = chunk_coord.r; q.chunk_c = chunk_coord.c; q.layer = layer; q.faction_id = faction_id;
    q.tgt_chunk_r = target.chunk_r; q.tgt_chunk_c = target.chunk_c; q.tgt_tile_r = target.tile_r; q.tgt_tile_c = target.tile_c;
    q.prev_index = -1;
    if(prev_los) {
        q.prev_index = PFNAV_LOS_PREV_INPLACE;
        q.prev_chunk_r = prev_los->chunk.r; q.prev_chunk_c = prev_los->chunk.c;
        memcpy(bytes, prev_los->field, sizeof(bytes));  /* one byte per tile: bit 0 visible, bit 1 wavefront_blocked */
    }
    SHIM_CHK(pfnav_los_fields_create(s->nav, &q, 1, bytes));
    out_los->chunk = chunk_coord;
    memcpy(out_los->field, bytes, sizeof(bytes));
    pthread_mutex_unlock(&s_lock);
}

/*****************************************************************************/
/* nav.h: the rest of field.c's exports                                      */
/*****************************************************************************/

vec2_t N_FlowDir(enum flow_dir dir)
{
    /* field.c:2429: unit vectors, x decreases with the column; the diagonal component is the float nearest 1/sqrt(2) */
    const float d = (float)(1.0 / 1.4142135623730951);
    switch(dir) {
    case FD_NW:   return (vec2_t){ d, -d};
    case FD_N:    return (vec2_t){ 0.0f, -1.0f};
    case FD_NE:   return (vec2_t){-d, -d};
    case FD_W:    return (vec2_t){ 1.0f, 0.0f};
    case FD_E:    return (vec2_t){-1.0f, 0.0f};
    case FD_SW:   return (vec2_t){ d,  d};
    case FD_S:    return (vec2_t){ 0.0f, 1.0f};
    case FD_SE:   return (vec2_t){-d,  d};
    default:      return (vec2_t){ 0.0f, 0.0f};
    }
}

static void shim_sync_region(struct shim_ctx *s, const struct nav_private *priv, enum nav_layer layer, struct tile_desc center, int dim)
{
    const int ar = center.chunk_r * FIELD_RES_R + center.tile_r, ac = center.chunk_c * FIELD_RES_C + center.tile_c;
    const int r0 = (ar - dim) / FIELD_RES_R - 1, r1 = (ar + dim) / FIELD_RES_R + 1;
    const int c0 = (ac - dim) / FIELD_RES_C - 1, c1 = (ac + dim) / FIELD_RES_C + 1;
    for(int r = r0; r <= r1; r++)
    for(int c = c0; c <= c1; c++) {
        if(r < 0 || c < 0 || r >= (int)priv->height || c >= (int)priv->width) continue;
        shim_sync_chunk(s, priv, layer, (struct coord){r, c});
    }
}

static void shim_region(void *nav_private, enum nav_layer layer, uint16_t enemies, size_t rdim, size_t cdim,
                        const int32_t *seeds_rc, size_t nseeds, struct tile_desc center, const struct nav_cell_overlay *overlay,
                        uint32_t flags, struct tile_desc start, uint8_t *inout)
{
    const struct nav_private *priv = nav_private;
    assert(rdim == cdim && rdim <= PFNAV_REGION_DIM_MAX);       /* the reference's own callers use squares (formation.c:3152) */
    pthread_mutex_lock(&s_lock);
    struct shim_ctx *s = shim_for(priv);
    shim_build_layer(s, priv, layer);
    shim_sync_region(s, priv, layer, center, (int)rdim);
    if(enemies) {
        const size_t nchunks = priv->width * priv->height;
        uint8_t *fac = malloc(nchunks * MAX_FACTIONS * 4096);
        for(size_t i = 0; i < nchunks; i++)
            memcpy(fac + i * MAX_FACTIONS * 4096, priv->chunks[layer][i].factions, MAX_FACTIONS * 4096);
        SHIM_CHK(pfnav_map_upload_factions(s->nav, layer, fac));
        free(fac);
    }
    pfnav_region_req q;
    memset(&q, 0, sizeof(q));
    q.layer = layer;
    q.center_r = center.chunk_r * FIELD_RES_R + center.tile_r; q.center_c = center.chunk_c * FIELD_RES_C + center.tile_c;
    q.start_r = start.chunk_r * FIELD_RES_R + start.tile_r; q.start_c = start.chunk_c * FIELD_RES_C + start.tile_c;
    q.seed_off = 0; q.seed_n = (int32_t)nseeds;
    q.overlay_off = 0; q.overlay_n = overlay ? (int32_t)overlay->nblocked : 0;
    q.enemies = enemies; q.flags = (uint16_t)flags;
    int32_t *ov = NULL;
    if(q.overlay_n) {
        ov = malloc(sizeof(int32_t) * 2 * q.overlay_n);
        for(int i = 0; i < q.overlay_n; i++) {
            ov[2*i] = overlay->blocked[i].chunk_r * FIELD_RES_R + overlay->blocked[i].tile_r;
            ov[2*i+1] = overlay->blocked[i].chunk_c * FIELD_RES_C + overlay->blocked[i].tile_c;
        }
    }
    SHIM_CHK(pfnav_region_fields(s->nav, (int)rdim, &q, 1, seeds_rc, nseeds, ov, q.overlay_n, inout));
    free(ov);
    pthread_mutex_unlock(&s_lock);
}

void N_CellArrivalFieldCreate(void *nav_private, size_t rdim, size_t cdim, enum nav_layer layer, uint16_t enemies,
                              struct tile_desc target, struct tile_desc center, uint8_t *out, void *workspace,
                              size_t workspace_size, const struct nav_cell_overlay *overlay)
{
    (void)workspace; (void)workspace_size;      /* the integration buffers live in shared memory on the device */
    const int32_t seed[2] = {target.chunk_r * FIELD_RES_R + target.tile_r, target.chunk_c * FIELD_RES_C + target.tile_c};
    shim_region(nav_private, layer, enemies, rdim, cdim, seed, 1, center, overlay, PFNAV_REGION_CREATE | PFNAV_REGION_CELL,
                (struct tile_desc){0, 0, 0, 0}, out);
}

void N_CellArrivalFieldUpdateToNearestPathable(void *nav_private, size_t rdim, size_t cdim, enum nav_layer layer,
                                               uint16_t enemies, struct tile_desc start, struct tile_desc center,
                                               uint8_t *inout, void *workspace, size_t workspace_size,
                                               const struct nav_cell_overlay *overlay)
{
    (void)workspace; (void)workspace_size;
    shim_region(nav_private, layer, enemies, rdim, cdim, NULL, 0, center, overlay, PFNAV_REGION_FIXUP, start, inout);
}

void N_GroupArrivalFieldCreate(void *nav_private, size_t rdim, size_t cdim, enum nav_layer layer, uint16_t enemies,
                               vec3_t map_pos, const vec2_t *targets, size_t ntargets, vec2_t center, uint8_t *out,
                               void *workspace, size_t workspace_size, const struct nav_cell_overlay *overlay)
{
    (void)workspace; (void)workspace_size;
    const struct nav_private *priv = nav_private;
    assert(rdim == cdim);
    pthread_mutex_lock(&s_lock);
    struct shim_ctx *s = shim_for(priv);
    shim_build_layer(s, priv, layer);
    if(s->map_x != map_pos.x || s->map_z != map_pos.z) {
        SHIM_CHK(pfnav_map_set_pos(s->nav, map_pos.x, map_pos.z));
        s->map_x = map_pos.x; s->map_z = map_pos.z;
    }
    for(size_t r = 0; r < priv->height; r++)        /* the centre is a world position: mirror the layer's occupancy */
    for(size_t c = 0; c < priv->width; c++)
        shim_sync_chunk(s, priv, layer, (struct coord){(int)r, (int)c});
    int32_t *ov = NULL;
    const size_t nov = overlay ? overlay->nblocked : 0;
    if(nov) {
        ov = malloc(sizeof(int32_t) * 2 * nov);
        for(size_t i = 0; i < nov; i++) {
            ov[2*i] = overlay->blocked[i].chunk_r * FIELD_RES_R + overlay->blocked[i].tile_r;
            ov[2*i+1] = overlay->blocked[i].chunk_c * FIELD_RES_C + overlay->blocked[i].tile_c;
        }
    }
    SHIM_CHK(pfnav_group_arrival_field(s->nav, layer, (int)rdim, enemies, (const float*)targets, ntargets,
                                       (const float*)&center, ov, nov, out));
    free(ov);
    pthread_mutex_unlock(&s_lock);
}
