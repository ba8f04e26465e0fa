/*
 * pfnav.h -- C ABI of libpfnav.so: the B200 (sm_100a) implementation of permafrost-engine's
 * per-tick navigation + crowd-movement hot path.
 *
 * Plain C, plain pointers and sizes; no CUDA or torch types in any signature (streams are passed
 * as an opaque `void*` that is a cudaStream_t / CUstream, NULL = the legacy default stream).
 * All functions return 0 on success and a negative pfnav_status on failure;
 * pfnav_last_error() returns a thread-local human-readable message.  There is NO CPU fallback:
 * if no sm_100-class device is usable, pfnav_create() fails.
 *
 * Each entry point cites the reference interface (file:line under the reference checkout) that
 * it replaces; INTEGRATION.md shows the binding a maintainer adds on the engine side.
 *
 * Conventions (identical to the reference):
 *   - a map is chunk_w x chunk_h chunks; a chunk is 64x64 nav tiles (FIELD_RES_R/C,
 *     src/navigation/nav_data.h:45-46); a nav tile is 4x4 world units; chunk = 256x256 wu.
 *   - world X DEcreases with the column index, Z increases with the row index
 *     (src/map/tile.c:547-592); map_pos is the (x,z) of the top-left corner.
 *   - chunk-blocked grids are [chunk_r][chunk_c][64][64], the order of N_CopyCostBasePacked
 *     (src/navigation/nav.c:2432).
 *   - a flow field is 4096 bytes, one `dir_idx` (enum flow_dir 0..8, nav.h:94-104) per tile,
 *     byte-identical to `struct flow_field::field` (src/navigation/field.h:103-109).
 *   - a LOS field is 4096 bytes, bit0 = visible, bit1 = wavefront_blocked, byte-identical to
 *     `struct LOS_field::field` (field.h:48-54).
 */
#ifndef PFNAV_H
#define PFNAV_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PFNAV_FIELD_RES        64
#define PFNAV_FIELD_TILES      4096
#define PFNAV_COST_IMPASSABLE  0xff     /* nav_data.h:47 */
#define PFNAV_ISLAND_NONE      0xffff   /* nav_data.h:48 */
#define PFNAV_FACTION_ID_NONE  0xf      /* nav_data.h:49 */
#define PFNAV_NAV_LAYER_MAX    12       /* nav.h:78-92 */
#define PFNAV_MAX_NEIGHBOURS   32       /* movement.c:437 */

typedef enum pfnav_status {
    PFNAV_OK            =  0,
    PFNAV_ERR_ARG       = -1,
    PFNAV_ERR_CUDA      = -2,
    PFNAV_ERR_NO_DEVICE = -3,
    PFNAV_ERR_STATE     = -4,
    PFNAV_ERR_NOMEM     = -5
} pfnav_status;

/* enum flow_dir, src/navigation/public/nav.h:94-104 */
enum pfnav_flow_dir { PFNAV_FD_NONE = 0, PFNAV_FD_NW, PFNAV_FD_N, PFNAV_FD_NE, PFNAV_FD_W,
                      PFNAV_FD_E, PFNAV_FD_SW, PFNAV_FD_S, PFNAV_FD_SE };

/* field_target.type, src/navigation/field.h:78-91 (only the chunk-local kinds) */
enum pfnav_target_type { PFNAV_TARGET_PORTAL = 0, PFNAV_TARGET_TILE = 1 };

typedef struct pfnav_ctx pfnav_ctx;

/* ---------------------------------------------------------------------------------------- */
/* Context + map state (replaces struct nav_private / nav_chunk, nav_private.h:52, nav_data.h:118)
 * ---------------------------------------------------------------------------------------- */

const char *pfnav_last_error(void);
int  pfnav_version(void);

/* device: CUDA ordinal. Fails (PFNAV_ERR_NO_DEVICE) without a usable GPU. */
/* Streams: every `void *stream` argument is a cudaStream_t. NULL names the context's own
 * non-blocking stream -- NOT the legacy default stream (work queued there would not be ordered
 * against the context's stream). Calls on one stream are ordered as issued. */
int  pfnav_create(int device, pfnav_ctx **out);
void pfnav_destroy(pfnav_ctx *ctx);

/* N_NewCtxForMapData (nav.c:2284): allocate device grids for chunk_w x chunk_h chunks and
 * `nlayers` navigation layers (cost_base u8, blockers u16, local_islands u16 per layer). */
int  pfnav_map_create(pfnav_ctx *ctx, int chunk_w, int chunk_h, int nlayers,
                      float map_x, float map_z);

/* map->pos: the world (x, z) of the map's top-left corner, when it is only known after the grids (a caller that
 * mirrors an existing nav context learns it with the first world-space call). Grids are unaffected. */
int  pfnav_map_set_pos(pfnav_ctx *ctx, float map_x, float map_z);

/* Upload one layer from HOST chunk-blocked arrays ([chunk_r][chunk_c][64][64]).
 * blockers / local_islands may be NULL (treated as all 0 / left unchanged).
 * Same packed layout as N_CopyCostBasePacked / N_CopyBlockersPacked (nav.c:2432, 2462). */
int  pfnav_map_upload_layer(pfnav_ctx *ctx, int layer, const uint8_t *cost_base,
                            const uint16_t *blockers, const uint16_t *local_islands);

/* Cost pass of N_NewCtxForMapData for one layer ON THE DEVICE (nav.c:2311-2336): n_set_cost_for_tile
 * (nav.c:267) for every tile type incl. corner tiles (M_Tile_{NW,NE,SW,SE}Height, map/tile.c:117-180)
 * + n_make_cliff_edges (nav.c:431); blockers of the layer are cleared (nav.c:2332). Bit-exact.
 *   chunk_tiles[chunk_r*chunk_w + chunk_c] -> 32x32 records of `tile_stride` bytes whose first 16
 *   bytes are the head of the engine's `struct tile` (map/public/tile.h:101): bool pathable @0,
 *   enum tiletype @4, int base_height @8, int ramp_height @12. Pass the engine's own
 *   `const struct tile **chunk_tiles` with tile_stride = sizeof(struct tile) (36).
 *   ref_layer: the reference's enum nav_layer (0-3 ground, 4-7 water, 8-11 air; nav.h:78-92) whose
 *   pathability rule applies to `layer`. Follow with pfnav_map_build_nav(layer). */
int  pfnav_map_cost_from_tiles(pfnav_ctx *ctx, int layer, int ref_layer, const void *const *chunk_tiles,
                               size_t tile_stride);
/* PFMAP terrain ingestion (docs/pfmap.txt). pfnav_tile is the head of the engine's `struct tile`
 * (map/public/tile.h:101) as pfnav_map_cost_from_tiles reads it (tile_stride = sizeof(pfnav_tile)).
 * pfnav_pfmap_parse: al_parse_pfmap_header (asset_load.c:168) + the material / splat lines + m_al_read_pfchunk /
 * m_al_parse_tile (map_asset_load.c:181, 103) over a memory image of the file. out_tiles = [chunk][32][32], chunks in
 * row-major order, cap_tiles records; NULL reads the header only. Malformed input -> PFNAV_ERR_ARG. Host only.
 * pfnav_map_load_pfmap: parse + pfnav_map_create(num_cols, num_rows, nlayers) + for every layer i the device cost
 * pass under the reference layer ref_layers[i] + pfnav_map_build_nav(i) -- N_NewCtxForMapData (nav.c:2284). */
typedef struct pfnav_tile { int32_t pathable, type, base_height, ramp_height; } pfnav_tile;
int  pfnav_pfmap_parse(const char *text, size_t len, int *out_chunk_rows, int *out_chunk_cols, pfnav_tile *out_tiles,
                       size_t cap_tiles);
int  pfnav_map_load_pfmap(pfnav_ctx *ctx, const char *text, size_t len, int nlayers, const int32_t *ref_layers,
                          float map_x, float map_z);
/* Read one layer's DEVICE grids back in the packed chunk-blocked layout of N_CopyCostBasePacked /
 * N_CopyBlockersPacked (nav.c:2432, 2462). Any out pointer may be NULL. */
int  pfnav_map_get_layer(pfnav_ctx *ctx, int layer, uint8_t *cost_base, uint16_t *blockers,
                         uint16_t *local_islands);

/* Sparse update of one chunk (what N_BlockersIncref/Decref + N_Update change, nav.c:4663-4705,
 * 2119): any pointer may be NULL to keep that grid. HOST pointers, [64][64] each. */
int  pfnav_map_update_chunk(pfnav_ctx *ctx, int layer, int chunk_r, int chunk_c,
                            const uint8_t *cost_base, const uint16_t *blockers,
                            const uint16_t *local_islands);

/* Structural part of N_NewCtxForMapData for one layer (nav.c:2284), computed on the host from the
 * uploaded cost_base + blockers: local islands (n_update_local_islands, nav.c:967: ids from 1 in
 * row-major discovery order) and the portal table (n_create_portals / n_link_chunks, nav.c:563,
 * 477: identical indices and endpoints). The islands are pushed to the device. */
int  pfnav_map_build_nav(pfnav_ctx *ctx, int layer);
/* n_update_dirty_local_islands (nav.c:986) for one chunk after pfnav_map_update_chunk changed
 * its blockers; pushes the refreshed islands to the device. */
int  pfnav_map_refresh_chunk(pfnav_ctx *ctx, int layer, int chunk_r, int chunk_c);
/* Read back: local islands ([chunk][64][64] u16) and portals (10 ints per portal: chunk_r, chunk_c,
 * index, endpoints[0].r/.c, endpoints[1].r/.c, connected chunk index, connected portal index, 0). */
int  pfnav_local_islands_get(pfnav_ctx *ctx, int layer, uint16_t *out);
int  pfnav_portals_get(pfnav_ctx *ctx, int layer, int32_t *out, int maxout, int *out_n);

/* "Attacking" requests (N_RequestPathAttacking, nav.c:3393): a field / LOS request whose faction_id is not
 * PFNAV_FACTION_ID_NONE treats a blocked tile as passable when every faction holding a blocker refcount on it
 * is at war with the requesting faction (field_tile_passable_no_enemies, field.c:179).
 *   pfnav_set_enemy_factions: G_GetEnemyFactions (game.h:184), bit i of the mask = faction i is an enemy.
 *   pfnav_map_upload_factions: chunk->factions of one layer, u8 [chunk][15][64][64]; pfnav_blockers_incref /
 *   decref maintain the same counts when given a faction id in 0..14 (nav.c:1032). */
int  pfnav_set_enemy_factions(pfnav_ctx *ctx, int faction_id, uint16_t enemies_mask);
int  pfnav_map_upload_factions(pfnav_ctx *ctx, int layer, const uint8_t *factions);
/* N_BlockersIncref / N_BlockersDecref (nav.c:4663-4683): reference-count the tiles under a circle
 * (M_Tile_AllUnderCircle, tile.c:687) plus 1/2/3 contour rings (M_Tile_Contour, tile.c:759) for the
 * 3x3/5x5/7x7 layers, on the ground AND water layers for non-air entities, air layers otherwise.
 * Layers the context does not hold are skipped. Host-side; the device sees the change at commit. */
int  pfnav_blockers_incref(pfnav_ctx *ctx, float x, float z, float range, int faction_id, uint32_t flags);
int  pfnav_blockers_decref(pfnav_ctx *ctx, float x, float z, float range, int faction_id, uint32_t flags);
/* N_BlockersIncrefOBB / N_BlockersDecrefOBB (nav.c:4685-4705): building footprints. corners_xz = the bottom
 * face of the engine's `struct obb`: corners[0], [1], [5], [4] as 4 (x, z) pairs (M_Tile_AllUnderObj,
 * tile.c:594-599). All four corners must lie inside the map. */
int  pfnav_blockers_incref_obb(pfnav_ctx *ctx, const float *corners_xz, int faction_id, uint32_t flags);
int  pfnav_blockers_decref_obb(pfnav_ctx *ctx, const float *corners_xz, int faction_id, uint32_t flags);
/* a tick's worth of the two calls above, applied in order (delta = +1 incref, -1 decref) */
typedef struct pfnav_blocker_op { float x, z, range; int32_t faction_id; uint32_t flags; int32_t delta; } pfnav_blocker_op;
int  pfnav_blockers_batch(pfnav_ctx *ctx, const pfnav_blocker_op *ops, size_t n);
int  pfnav_blockers_get(pfnav_ctx *ctx, int layer, uint16_t *out);
/* nav_chunk::factions of one layer: out = u8 [chunks][15][64][64] */
int  pfnav_blockers_get_factions(pfnav_ctx *ctx, int layer, uint8_t *out);
/* N_Update + N_ApplyDeferredInvalidations (nav.c:2119-2223): for every chunk whose occupancy
 * changed, recompute the local islands, refresh the portal edge states, push the chunk's blockers +
 * islands to the device, and invalidate the pool: entries AT the chunk, and -- when an edge state
 * flipped -- every field of every destination routed THROUGH it (fieldcache.c:460-545). */
int  pfnav_map_commit(pfnav_ctx *ctx, int *out_ndirty);

/* ---------------------------------------------------------------------------------------- */
/* Flow fields + LOS fields (seam B2: src/navigation/field.h:115-202)
 * ---------------------------------------------------------------------------------------- */

/* One N_FlowFieldUpdate call (field.c:2030). 64 bytes. Mirrors struct field_work_in
 * (nav.c:126-131) with the portal_desc pointers resolved to plain coordinates. */
typedef struct pfnav_field_req {
    int32_t  chunk_r, chunk_c;      /* struct coord chunk_coord */
    int32_t  layer;                 /* enum nav_layer */
    int32_t  faction_id;            /* PFNAV_FACTION_ID_NONE for non-attacking paths */
    int32_t  target_type;           /* enum pfnav_target_type */
    int32_t  init;                  /* !=0: N_FlowFieldInit first (field.c:2020); 0: update the
                                     * existing `inout` field in place (nav.c:1998-2008) */
    /* TARGET_TILE: struct coord tile */
    int32_t  tile_r, tile_c;
    /* TARGET_PORTAL: struct portal_desc (field.h:68-73) */
    int16_t  port_r0, port_c0, port_r1, port_c1;   /* pd.port->endpoints[0..1] */
    int16_t  next_r0, next_c0, next_r1, next_c1;   /* pd.next->endpoints[0..1] */
    int32_t  next_chunk_r, next_chunk_c;           /* pd.next->chunk (== port->connected chunk) */
    uint16_t port_iid, next_iid;                   /* pd.port_iid, pd.next_iid */
    int32_t  _pad;
} pfnav_field_req;

/* Batch of N_FlowFieldUpdate (+ optional N_FlowFieldInit). reqs and inout_fields are HOST
 * buffers (inout_fields: n x 4096 bytes). Synchronous: results are in inout_fields on return. */
int  pfnav_flow_fields_update(pfnav_ctx *ctx, const pfnav_field_req *reqs, size_t n,
                              uint8_t *inout_fields);

/* Repair chain of N_DesiredPointSeekVelocity (nav.c:3508-3554) for an entity standing on a tile whose
 * cached direction is FD_NONE, applied in place to HOST fields (4096 bytes each):
 *   PFNAV_REPAIR_NEAREST_PATHABLE  N_FlowFieldUpdateToNearestPathable (field.c:2247): args[i] =
 *       start_r << 8 | start_c, a NON-passable tile of the chunk; directions lead out of the blocked blob
 *   PFNAV_REPAIR_ISLAND_TO_NEAREST N_FlowFieldUpdateIslandToNearest (field.c:2307): args[i] = the local
 *       island id cut off from the field's frontier; needs pfnav_route_build(layer) (global islands)
 * targets[i] = the request that built field i (its chunk, layer and target). */
enum pfnav_repair_kind { PFNAV_REPAIR_NEAREST_PATHABLE = 0, PFNAV_REPAIR_ISLAND_TO_NEAREST = 1 };
int  pfnav_flow_fields_repair(pfnav_ctx *ctx, const pfnav_field_req *targets, const int32_t *kinds,
                              const int32_t *args, size_t n, uint8_t *inout_fields);

/* Same, DEVICE-resident reqs / fields, asynchronous on `stream`. */
int  pfnav_flow_fields_update_dev(pfnav_ctx *ctx, const pfnav_field_req *d_reqs, size_t n,
                                  uint8_t *d_inout_fields, void *stream);

/* One N_LOSFieldCreate call (field.c:2085). 48 bytes. */
#define PFNAV_LOS_PREV_INPLACE (-3)
typedef struct pfnav_los_req {
    int32_t  chunk_r, chunk_c;          /* chunk the field is for */
    int32_t  layer;
    int32_t  faction_id;
    int32_t  tgt_chunk_r, tgt_chunk_c;  /* struct tile_desc target */
    int32_t  tgt_tile_r, tgt_tile_c;
    int32_t  prev_index;                /* -1: destination chunk (prev_los == NULL); >= 0: index,
                                         * within the same batch, of the request whose output is
                                         * prev_los (must be < this request's own index);
                                         * PFNAV_LOS_PREV_INPLACE (host API): prev_los is a field the
                                         * caller holds -- its 4096 bytes are passed in out_fields[i]
                                         * and replaced by the result (N_LOSFieldCreate's `prev`).  */
    int32_t  prev_chunk_r, prev_chunk_c;/* prev_los->chunk */
    int32_t  _pad;
} pfnav_los_req;

/* Batch of N_LOSFieldCreate. Requests are processed in dependency order (a request may name an
 * earlier one as its prev_los). HOST buffers; out_fields: n x 4096 bytes. */
int  pfnav_los_fields_create(pfnav_ctx *ctx, const pfnav_los_req *reqs, size_t n,
                             uint8_t *out_fields);
int  pfnav_los_fields_create_dev(pfnav_ctx *ctx, const pfnav_los_req *d_reqs, size_t n,
                                 uint8_t *d_out_fields, int n_waves, const int32_t *h_wave_offsets,
                                 void *stream);

/* Goal planner (the request-generation half of n_request_path, nav.c:1774-2047): emits, for every
 * chunk connected to the goal tile, the TARGET_TILE / TARGET_PORTAL flow request and the chained
 * LOS request that guide agents there. Breadth-first over (chunk, local island) nodes, i.e. it
 * minimises portal hops rather than AStar_PortalGraphPath's travel cost (a_star.c:429).
 *   flow_chunk[i] / los_chunk[i] : chunk index (r*chunk_w+c) a request's output belongs to
 *   flow_wave[i]                 : 0 for a chunk's first request (init=1); k>0 for the k-th further
 *                                  target merged into the same field (init=0, nav.c:1998-2008);
 *                                  wave k must run after wave k-1 on that chunk's field.
 *   los requests are emitted parent-before-child with prev_index into los_out. */
int  pfnav_plan_goal(pfnav_ctx *ctx, int layer, int tgt_chunk_r, int tgt_chunk_c, int tgt_tile_r,
                     int tgt_tile_c, pfnav_field_req *flow_out, int32_t *flow_chunk,
                     int32_t *flow_wave, int max_flow, int *n_flow, pfnav_los_req *los_out,
                     int32_t *los_chunk, int max_los, int *n_los);

/* Cost-faithful routing structure of one layer (after pfnav_map_build_nav): portal edges with
 * AStar_GridPath costs (nav.c:593, a_star.c:303), per-portal travel index (nav.c:1314), global
 * islands (nav.c:1731), edge states (nav.c:693). Host-side; bit-identical costs and ids. */
int  pfnav_route_build(pfnav_ctx *ctx, int layer);
int  pfnav_route_islands_get(pfnav_ctx *ctx, int layer, uint16_t *out);
/* edges of one portal, 3 u32 per edge {neighbour portal_ref, edge_state, cost float bits} */
int  pfnav_route_edges_get(pfnav_ctx *ctx, int layer, int chunk, int portal, uint32_t *out, int maxout,
                           int *out_n);
/* n_request_path (nav.c:1774-2047), request-generation half: the flow requests (with their
 * N_FlowFieldID, field.c:1952) and LOS requests, in the reference's order, that the path src -> dst
 * needs given what (dest, chunk) entries already exist: have_flow[chunk] = mapped ff_id or 0,
 * have_los[chunk] = 0/1. LOS prev_index: -1 destination chunk, >= 0 index in this batch, -2 the
 * previous chunk's field pre-exists (caller supplies it). *out_ok = the function's bool result. */
int  pfnav_route_request_path(pfnav_ctx *ctx, int layer, float src_x, float src_z, float dst_x,
                              float dst_z, const uint64_t *have_flow, const uint8_t *have_los,
                              pfnav_field_req *flow_out, uint64_t *flow_ffid, int32_t *flow_chunk,
                              int max_flow, int *n_flow, pfnav_los_req *los_out, int32_t *los_chunk,
                              int max_los, int *n_los, uint32_t *out_dest_id, int *out_ok);
/* N_RequestPathAttacking (nav.c:3393): the faction applied to the path requests that follow (route + pool
 * entry points); PFNAV_FACTION_ID_NONE restores N_RequestPath. It is packed into the dest_id (nav.c:853) and
 * forwarded to every flow / LOS request, which is all n_request_path does with it. */
int  pfnav_request_faction(pfnav_ctx *ctx, int faction_id);
/* N_RequestPath (nav.c:3386) against the device field pool: route src -> dst exactly as the
 * reference does and build, on the device, the fields the pool does not hold yet for `dest`
 * (cached (dest, chunk) entries and their ff_ids are honoured like the field cache's). */
int  pfnav_pool_request_path(pfnav_ctx *ctx, int dest, int layer, float src_x, float src_z,
                             float dst_x, float dst_z, void *stream, uint32_t *out_dest_id,
                             int *out_ok, int *out_n_flow, int *out_n_los);
/* Read one pool entry back (4096 B each, pointers may be NULL). *out_has: bit0 flow, bit1 LOS;
 * *out_ffid: the N_FlowFieldID currently mapped for (dest, chunk). */
int  pfnav_pool_get(pfnav_ctx *ctx, int dest, int chunk_r, int chunk_c, uint8_t *flow_out,
                    uint8_t *los_out, int *out_has, uint64_t *out_ffid);
/* Host-only context for the HOST-side structure code above (islands, portals, routing) on a
 * machine without a GPU. It has no compute path: kernel-launching entry points fail with
 * PFNAV_ERR_NO_DEVICE. */
int  pfnav_create_hostonly(pfnav_ctx **out);

/* Selects how the 64x64 tiles are staged into shared memory: 1 = TMA tensor maps
 * (cp.async.bulk.tensor), 0 = plain coalesced loads. Default 1. */
int  pfnav_set_tma(pfnav_ctx *ctx, int enable);

/* ---------------------------------------------------------------------------------------- */
/* Field pool: device-resident (dest_id, chunk) -> {flow field, LOS field}
 * (replaces the fieldcache lookups done per agent: fieldcache.c N_FC_GetDestFFMapping /
 *  N_FC_FlowFieldAt / N_FC_LOSFieldAt, called from nav.c:3483-3506, 4044-4052)
 * ---------------------------------------------------------------------------------------- */

/* Size the pool for `ndests` destinations (flocks). Slot table is ndests x chunks. */
int  pfnav_pool_create(pfnav_ctx *ctx, int ndests, int max_fields);
/* Store field bytes (HOST, 4096 each; either may be NULL) for (dest, chunk). */
int  pfnav_pool_put(pfnav_ctx *ctx, int dest, int chunk_r, int chunk_c,
                    const uint8_t *flow_field, const uint8_t *los_field);
int  pfnav_pool_clear(pfnav_ctx *ctx);
/* N_RequestPath's field-building half (nav.c:1819-2042), resident on the device: plan the goal
 * (pfnav_plan_goal), then run the flow waves and the LOS chain straight into the pool slots of
 * destination `dest`. No field bytes cross PCIe. Launches are asynchronous on `stream`. */
int  pfnav_pool_request_goal(pfnav_ctx *ctx, int dest, int layer, int tgt_chunk_r, int tgt_chunk_c,
                             int tgt_tile_r, int tgt_tile_c, void *stream, int *out_n_flow,
                             int *out_n_los);
/* The same for a batch of goals (the reference runs up to 256 field tasks concurrently, nav.c:88):
 * the flow waves and the LOS dependency waves of all goals are launched together.
 * dests[ngoals]; targets: 4 ints per goal {chunk_r, chunk_c, tile_r, tile_c}. */
int  pfnav_pool_request_goals(pfnav_ctx *ctx, int ngoals, const int32_t *dests, int layer,
                              const int32_t *targets, void *stream, int *out_n_flow, int *out_n_los);
/* PFNAV_REQUEST_MISSING_ONLY: (dest, chunk) fields the pool still holds are kept -- the field cache's hit path
 * (fieldcache.c:138-141) -- and only the missing ones (never built, invalidated by pfnav_map_commit, or evicted)
 * are planned and built; a rebuilt LOS field reads a kept parent chunk's field out of the pool. */
#define PFNAV_REQUEST_MISSING_ONLY (1u << 0)
int  pfnav_pool_request_goals_ex(pfnav_ctx *ctx, int ngoals, const int32_t *dests, int layer, const int32_t *targets,
                                 uint32_t flags, void *stream, int *out_n_flow, int *out_n_los);
/* pfnav_pool_request_goals runs the LOS dependency chains on a context-owned stream so that work which
 * does not read fields (position index, cohesion) overlaps them; pfnav_agents_tick and the pool entry
 * points order themselves after it. A caller that reads pool LOS fields through raw device pointers
 * on its own stream calls this first. */
/* Test / tuning hook: LOS kernel variant, 0 = bit rows, 1 (default) = one state byte per tile. Same results. */
int  pfnav_set_los_variant(pfnav_ctx *ctx, int variant);
/* Profiling aid: per-field trace of the LOS launches. enable != 0 arms it for the following launches; with
 * out != NULL the trace of the last traced launch is returned first: out[4i..4i+3] = {taken, dependency
 * satisfied, finished} in %globaltimer nanoseconds and 1 + heap pops (0 = zero-filled early out). */
int  pfnav_los_trace(pfnav_ctx *ctx, int enable, unsigned long long *out, size_t cap, size_t *out_n);
int  pfnav_fields_join(pfnav_ctx *ctx, void *stream);
/* On-miss chain of N_DesiredPointSeekVelocity (nav.c:3484-3554) for the uploaded work list, against the pool:
 * every work agent whose own tile has no direction (field absent or FD_NONE) is collected on the device; per
 * (dest, chunk, local island | blocked tile), in work order, the host then (1) requests the path from that
 * agent's position (n_request_path, nav.c:3486/3499 == pfnav_pool_request_path), and if the tile is still
 * FD_NONE (2) repairs the field in place: N_FlowFieldUpdateToNearestPathable when the tile is non-passable,
 * N_FlowFieldUpdateIslandToNearest otherwise. Blocking; needs pfnav_route_build. The caller re-runs the tick. */
int  pfnav_pool_repair(pfnav_ctx *ctx, int *out_nrequests, int *out_nrepairs);

/* ---------------------------------------------------------------------------------------- */
/* Region fields: the dim x dim "cell arrival" fields a formation builds for every cell and
 * every arriving group (formation.c:3152-3176 builds CELL_ARRIVAL_FIELD_RES = 96 squares per cell). The region
 * is centred on `center`, may straddle chunks and hang over the map edge, and its directions are packed two
 * per byte (set_flow_cell, field.c:790: even column -> high nibble); a field is dim * dim / 2 bytes, row-major.
 * All tile coordinates are ABSOLUTE nav tiles (chunk * 64 + tile) as (r, c) int32 pairs.
 *   PFNAV_REGION_CREATE  N_CellArrivalFieldCreate (field.c:2445) / N_GroupArrivalFieldCreate (field.c:2525):
 *                        zero the field, seed seeds[seed_off .. seed_off + seed_n) (those inside the region) at
 *                        cost 0, Dijkstra over passable tiles (enemy-mask rule when enemies != 0; overlay tiles
 *                        are never entered), derive directions.
 *   PFNAV_REGION_CELL    with CREATE: one seed (the cell's tile); the region base shifts when the tile lies one
 *                        past the far edge (field.c:2477-2482). seed_n must be 1.
 *   PFNAV_REGION_FIXUP   N_CellArrivalFieldUpdateToNearestPathable (field.c:2603), after CREATE when both are
 *                        set (cell_field_fixup_task, formation.c:3171): directions off the blocked island that
 *                        holds (start_r, start_c), a NON-passable tile inside the map-clamped region.
 * Without PFNAV_REGION_CREATE the field is updated in place (the caller supplies it). */
enum { PFNAV_REGION_CREATE = 1, PFNAV_REGION_FIXUP = 2, PFNAV_REGION_CELL = 4 };
#define PFNAV_REGION_DIM_MAX 128
typedef struct pfnav_region_req {      /* 40 bytes */
    int32_t  layer;
    int32_t  center_r, center_c;
    int32_t  start_r, start_c;         /* PFNAV_REGION_FIXUP only */
    int32_t  seed_off, seed_n;         /* pairs, into seeds_rc */
    int32_t  overlay_off, overlay_n;   /* pairs, into overlay_rc (struct nav_cell_overlay, nav.h:695) */
    uint16_t enemies;                  /* enemy faction bit mask, 0 = plain passability */
    uint16_t flags;
} pfnav_region_req;
/* HOST buffers; synchronous. inout_fields: n * dim * dim / 2 bytes. dim even, <= PFNAV_REGION_DIM_MAX. */
int  pfnav_region_fields(pfnav_ctx *ctx, int dim, const pfnav_region_req *reqs, size_t n, const int32_t *seeds_rc,
                         size_t nseeds, const int32_t *overlay_rc, size_t noverlay, uint8_t *inout_fields);
/* DEVICE buffers (requests already validated by the caller), asynchronous on `stream`. */
int  pfnav_region_fields_dev(pfnav_ctx *ctx, int dim, const pfnav_region_req *d_reqs, size_t n,
                             const int32_t *d_seeds_rc, const int32_t *d_overlay_rc, uint8_t *d_inout_fields,
                             void *stream);
/* N_GroupArrivalFieldCreate (field.c:2525) with its world-space arguments: targets_xz[ntargets][2], center_xz[2].
 * A centre outside the map yields the zero field, targets outside the map or the region are skipped. */
int  pfnav_group_arrival_field(pfnav_ctx *ctx, int layer, int dim, uint16_t enemies, const float *targets_xz,
                               size_t ntargets, const float *center_xz, const int32_t *overlay_rc, size_t noverlay,
                               uint8_t *out_field);

/* Group arrival (TARGET_ZONE) fields: a flock arriving at its goal follows one shared field per chunk into the
 * open tiles around the goal (arrival.c:859 -> N_RequestAsyncGroupArrivalField nav.c:3921 -> field task ->
 * N_FlowFieldUpdate field.c:2050 -> field_update_zone :1810). The seeds are the first pi * radius^2 open tiles of a
 * best-first flood from the centre (field_zone_initial_frontier, field.c:1683; computed on the host with the
 * reference's heap order), the integration runs on the device over the chunk padded by half a chunk, and the
 * chunk's 64 x 64 window is written one direction per byte. Call pfnav_map_commit first when blockers changed.
 * Only maps with more than one chunk row AND column (or exactly 1 x 1) are accepted: with one of the two the
 * reference indexes its padded buffer out of bounds. centre_r / centre_c: absolute nav tile of the zone centre. */
int  pfnav_zone_seeds(pfnav_ctx *ctx, int layer, int chunk_r, int chunk_c, int centre_r, int centre_c, int radius,
                      int32_t *out_rc, size_t cap, size_t *out_n);
/* chunks_rc: n (chunk_r, chunk_c) pairs; out_fields: HOST, n * 4096 bytes. Synchronous. */
int  pfnav_zone_fields(pfnav_ctx *ctx, int layer, int centre_r, int centre_c, int radius, const int32_t *chunks_rc,
                       size_t n, uint8_t *out_fields);
/* N_RequestAsyncGroupArrivalField + N_AwaitAsyncFields (nav.c:3921, 3958): the zone fields of every chunk within
 * reach (2 * radius tiles) of the world-space centre go into pool destination `dest` as its flow fields. A centre
 * outside the map requests nothing. Returns when the fields are in the pool. */
int  pfnav_pool_request_zone(pfnav_ctx *ctx, int dest, int layer, const float *centre_xz, int radius, void *stream,
                             int *out_nfields);
/* N_DesiredGroupArrivalVelocity (nav.c:3561) for n world positions against pool destination `dest`:
 * out_vel[n][2] = N_FlowDir of the tile's direction; out_flags[n]: bit 0 = the call returned true (position and centre on
 * the map, chunk field present), bit 1 = *out_at_slot (FD_NONE within `radius` tiles of the centre). HOST buffers. */
int  pfnav_group_arrival_velocity(pfnav_ctx *ctx, int dest, const float *centre_xz, int radius, const float *pos_xz,
                                  size_t n, float *out_vel, uint8_t *out_flags);

/* TARGET_ENTITY / TARGET_ENEMIES chunk fields (N_FlowFieldUpdate field.c:2040-2048 -> field_update_entity :1609,
 * field_update_enemies :1540; consumers N_DesiredSurroundVelocity / N_DesiredEnemySeekVelocity nav.c:3683, 3603).
 * Same padded-chunk integration as the zone fields; the zero-cost frontier is the set of tiles under the target
 * entity (field_entity_initial_frontier, field.c:1317) or under every enemy near the chunk
 * (field_enemies_initial_frontier, field.c:1209). What field.c reads about an entity through its
 * nav_unit_query_ctx is passed in a pfnav_footprint; WHICH entities count as enemies (faction, COMBATABLE,
 * diplomacy, not dying: field_enemy_ent field.c:963) is engine state and is decided by the caller. */
typedef struct pfnav_footprint {       /* 48 bytes */
    float    x, z;                     /* ent_pos_xz */
    float    sel_radius;               /* ent_sel_radius; used unless is_building */
    uint32_t is_building;              /* ENTITY_FLAG_BUILDING: the footprint is the bottom face of ent_curr_obb ... */
    float    corners_xz[8];            /* ... as 4 (x, z) pairs, like pfnav_blockers_incref_obb; all inside the map */
} pfnav_footprint;
enum { PFNAV_TARGET_ENTITY = 0, PFNAV_TARGET_ENEMIES = 1 };
/* Seed tiles of one chunk's field (host structure code, no device needed). ref_layer = the reference's enum
 * nav_layer of `layer`, which fixes the number of contour rings (ENTITY: 3x3 -> 1, 5x5 -> 1, 7x7 and above -> 2;
 * ENEMIES: 1 / 2 / 3). ENEMIES keeps only the entities whose position lies in the chunk's search rectangle
 * (the chunk grown by half a chunk + 16 units, field.c:1229-1243) and returns each tile once, row-major. */
int  pfnav_entity_seeds(pfnav_ctx *ctx, int ref_layer, int target_kind, const pfnav_footprint *ents, size_t nents,
                        int chunk_r, int chunk_c, int32_t *out_rc, size_t cap, size_t *out_n);
/* N_FlowFieldInit + N_FlowFieldUpdate for the listed chunks; out_fields: HOST, n * 4096 bytes. Synchronous. */
int  pfnav_entity_fields(pfnav_ctx *ctx, int layer, int ref_layer, int target_kind, const pfnav_footprint *ents,
                         size_t nents, const int32_t *chunks_rc, size_t n, uint8_t *out_fields);

/* ---------------------------------------------------------------------------------------- */
/* Per-tick agent velocity update (seam B1: the reference's own GPU back-end calls
 * R_GL_MoveUploadData / R_GL_MoveDispatchWork / R_GL_MoveReadNewVelocities,
 * src/render/public/render.h:619-691, computing what move_velocity_work computes on the CPU,
 * src/game/movement.c:3395-3466)
 * ---------------------------------------------------------------------------------------- */

/* enum move_state, movement.c:117-142 */
enum pfnav_move_state { PFNAV_STATE_MOVING = 0, PFNAV_STATE_MOVING_IN_FORMATION, PFNAV_STATE_ARRIVED,
                        PFNAV_STATE_SEEK_ENEMIES, PFNAV_STATE_WAITING, PFNAV_STATE_SURROUND_ENTITY,
                        PFNAV_STATE_ENTER_ENTITY_RANGE, PFNAV_STATE_TURNING, PFNAV_STATE_ARRIVING_TO_CELL };

/* Entity flags the path reads: the engine's own bit values (src/entity.h:59-82), so the
 * entity flag word can be passed through verbatim. Only the low 24 bits are kept. */
#define PFNAV_FLAG_MOVABLE      (1u << 3)
#define PFNAV_FLAG_WATER        (1u << 14)
#define PFNAV_FLAG_AIR          (1u << 15)
#define PFNAV_FLAG_GARRISONED   (1u << 18)
#define PFNAV_FLAG_COMBAT_HELD  (1u << 21)

/* One entity as the movement tick sees it: struct gpu_ent_desc (movement.c:350-369) plus the
 * two things the CPU path uses and the GLSL record lacks (prev_pos). 64 bytes. */
typedef struct pfnav_agent {
    float    pos[2];        /* xz, gamestate snapshot position            */
    float    prev_pos[2];   /* ms->prev_pos: ClearPath self position + LOS sample (movement.c:4351,4137) */
    float    velocity[2];   /* ms->velocity, world units per tick          */
    float    vdes[2];       /* in: ignored when the tick computes vdes from the field pool */
    float    radius;        /* selection radius                            */
    float    max_speed;     /* ms->max_speed (wu/s)                        */
    float    speed;         /* entity_speed() (wu/s)                       */
    uint32_t state;         /* enum pfnav_move_state                       */
    uint32_t flags;         /* PFNAV_FLAG_*                                */
    int32_t  flock;         /* index into flocks, -1 = none                */
    uint32_t has_dest_los;  /* in: ignored when computed from the pool     */
    uint32_t aux_dest1;     /* 1 + field-pool destination of this entity's TARGET_ENEMIES fields (STATE_SEEK_ENEMIES,
                             * M_NavDesiredEnemySeekVelocity) or TARGET_ENTITY fields (STATE_SURROUND_ENTITY,
                             * M_NavDesiredSurroundVelocity), built with pfnav_pool_request_entity_fields; 0 = none */
} pfnav_agent;

typedef struct pfnav_flock {
    float    target[2];     /* flock->target_xz (movement.c:193)           */
    int32_t  dest;          /* field-pool destination index, -1 = none     */
    int32_t  layer;         /* N_DestLayer(dest_id)                        */
} pfnav_flock;

/* move_copy_gamestate (movement.c:3607): upload the entity snapshot. HOST pointers.
 * uid == index. Rebuilds the device spatial index (position.c:359 G_Pos_CopyBitmapGrid). */
int  pfnav_agents_upload(pfnav_ctx *ctx, const pfnav_agent *agents, size_t n,
                         const pfnav_flock *flocks, size_t nflocks, int hz);

/* ---------------------------------------------------------------------------------------- */
/* Multi-GPU. Entities are partitioned by contiguous index range -- the reference's own fork-join
 * split of the work array (movement.c:3751-3762) -- one range per GPU; each GPU updates its own range and needs the
 * 24-byte neighbour record {pos, vel, radius, state | flags} of every other entity: ONE all-gather of those records
 * per tick, nothing else. The map is replicated; a rank's field pool holds the destinations of its own flocks.
 *
 *   one process per GPU (NCCL over NVLink / NVSwitch; libnccl.so.2 is resolved with dlopen at run time):
 *       rank 0: pfnav_mgpu_unique_id(id) -> the host program hands the 128 bytes to every rank
 *       all   : pfnav_mgpu_init(ctx, rank, world, id)
 *       upload: pfnav_agents_upload_shard(ctx, own records, lo, hi, n_total, ...)   [lo, hi) = pfnav_mgpu_shard_range
 *       tick  : pfnav_mgpu_gather(ctx, stream); pfnav_agents_tick(...); [compute / apply updates]; repeat
 *   one process driving several contexts from one thread (how the engine itself would use 8 GPUs):
 *       pfnav_group_create(ctxs, n, &g); per context pfnav_agents_upload_shard; pfnav_group_gather(g); ticks ...
 *
 * uids stay global (index into the whole population); work lists, movestate uploads, velocity / patch / state
 * read-backs of a context cover its OWN range only. */
typedef struct pfnav_group pfnav_group;
#define PFNAV_MGPU_ID_BYTES 128
int  pfnav_mgpu_shard_range(size_t n_total, int rank, int world, size_t *lo, size_t *hi);
int  pfnav_mgpu_unique_id(void *out_id /* PFNAV_MGPU_ID_BYTES */);
int  pfnav_mgpu_init(pfnav_ctx *ctx, int rank, int world, const void *id);
int  pfnav_mgpu_finalize(pfnav_ctx *ctx);
/* The per-tick collective: all-gather of the neighbour records (plus the flock id column after an upload that may
 * have changed membership), then the spatial index over the whole population (G_Pos_CopyBitmapGrid, position.c:359)
 * on this rank. Asynchronous on `stream` (blocks only when member lists are rebuilt). */
int  pfnav_mgpu_gather(pfnav_ctx *ctx, void *stream);
int  pfnav_group_create(pfnav_ctx **ctxs, int world, pfnav_group **out);
int  pfnav_group_gather(pfnav_group *g);       /* on every member's own context stream */
void pfnav_group_destroy(pfnav_group *g);
/* move_copy_gamestate (movement.c:3607) for this context's own range [lo, hi) of a population of n_total entities;
 * shard[i] is entity lo + i, `flock` fields index the GLOBAL flock table. The snapshot is complete after the next
 * pfnav_mgpu_gather / pfnav_group_gather. On a context outside any multi-GPU job lo = 0, hi = n_total is the whole
 * population (== pfnav_agents_upload).
 * PFNAV_UPLOAD_SAME_FLOCKS: only positions, velocities, states and speeds changed since the previous upload of the
 * same range -- flock membership, selection radii, layers, the GARRISONED flag and the flock table are as before --
 * so the per-entity host pass and the member-list rebuild are skipped (the engine pushes deltas). */
#define PFNAV_UPLOAD_SAME_FLOCKS (1u << 0)
int  pfnav_agents_upload_shard(pfnav_ctx *ctx, const pfnav_agent *shard, size_t lo, size_t hi, size_t n_total,
                               const pfnav_flock *flocks, size_t nflocks, int hz, uint32_t flags);

/* Work list (move_push_work, movement.c:3741): the uids that take a velocity update this tick.
 * NULL = every agent whose state is not ARRIVED/WAITING (ent_still, movement.c:652). */
int  pfnav_agents_set_work(pfnav_ctx *ctx, const uint32_t *uids, size_t nwork);

/* ---------------------------------------------------------------------------------------- */
/* State update: entity_compute_update (game/movement.c:2303-2650) on the device, one
 * `struct movestate_patch` (movement.c:245-262) per work item, for every movement state (the formation inputs
 * and the state beyond point seeking arrive through pfnav_formation_in / pfnav_movestate_ext). The engine applies the patches on
 * its main thread (entity_apply_update, movement.c:2693: blockers, events, G_Pos_Set);
 * pfnav_agents_apply_updates is the device-side equivalent for the movestate fields so that
 * consecutive ticks can run without a host round trip. */
#define PFNAV_VEL_HIST_LEN 14           /* VEL_HIST_LEN, movement.c:94 */
typedef struct pfnav_movestate {        /* the part of struct movestate (movement.c:146-215) the update reads */
    float   next_pos[3];                /* ms->next_pos                                  */
    float   step;                       /* ms->step                                      */
    float   next_rot[4];                /* ms->next_rot, quaternion {x, y, z, w}          */
    float   combat_facing[4];           /* ms->combat_facing                             */
    float   vel_hist[PFNAV_VEL_HIST_LEN][2];
    int32_t left;                       /* ms->left                                      */
    int32_t vel_hist_idx;
    int32_t _pad[2];
} pfnav_movestate;                      /* 176 bytes */

/* The formation inputs of one work item: struct formation_state + cell_pos + cell_arrival_vdes of struct move_work_in
 * (movement.c:215-225, 264-276). formation.c computes them on the main thread (G_Formation_*); the velocity pass and
 * the state update read them (STATE_MOVING_IN_FORMATION / STATE_ARRIVING_TO_CELL). 64 bytes. */
#define PFNAV_FORM_HAS_FORMATION    (1u << 0)   /* fstate.fid != NULL_FID */
#define PFNAV_FORM_ASSIGNMENT_READY (1u << 1)
#define PFNAV_FORM_ASSIGNED_TO_CELL (1u << 2)
#define PFNAV_FORM_IN_RANGE_OF_CELL (1u << 3)
#define PFNAV_FORM_ARRIVED_AT_CELL  (1u << 4)
typedef struct pfnav_formation_in {
    float    cell_pos[2];            /* in->cell_pos                                   */
    float    cell_arrival_vdes[2];   /* in->cell_arrival_vdes: the desired velocity of STATE_ARRIVING_TO_CELL */
    float    cohesion[2];            /* fstate.normal_cohesion_force                   */
    float    align[2];               /* fstate.normal_align_force                      */
    float    drag[2];                /* fstate.normal_drag_force                       */
    float    target_orientation[4];  /* fstate.target_orientation                      */
    uint32_t flags;                  /* PFNAV_FORM_*                                   */
    uint32_t _pad;
} pfnav_formation_in;
/* HOST array, one record per entity of this context's own range (uid order). Optional: without it every entity is
 * outside any formation (fid == NULL_FID). */
int  pfnav_agents_upload_formation(pfnav_ctx *ctx, const pfnav_formation_in *f, size_t n);

/* The rest of struct movestate (movement.c:146-215) that the states beyond point seeking read. 64 bytes. */
#define PFNAV_NULL_UID 0xffffffffu
typedef struct pfnav_movestate_ext {
    int32_t  wait_prev;              /* ms->wait_prev (enum pfnav_move_state)          */
    int32_t  wait_ticks_left;
    uint32_t surround_target_uid;    /* PFNAV_NULL_UID = none; also the target of STATE_ENTER_ENTITY_RANGE */
    uint32_t using_surround_field;
    float    target_range;           /* STATE_ENTER_ENTITY_RANGE                       */
    float    target_prev_pos[2];
    float    _padf;
    float    target_dir[4];          /* STATE_TURNING: ms->target_dir                  */
    float    rot[4];                 /* Entity_GetRot(uid), what Entity_SetRot last set */
} pfnav_movestate_ext;
/* HOST array, own range, uid order. Optional: zero / NULL_UID defaults. */
int  pfnav_agents_upload_movestate_ext(pfnav_ctx *ctx, const pfnav_movestate_ext *ms, size_t n);

/* enum movestate_flags (movement.c:226-242), same bit values */
#define PFNAV_UPDATE_SET_STATE        (1u << 0)
#define PFNAV_UPDATE_SET_VELOCITY     (1u << 1)
#define PFNAV_UPDATE_SET_POSITION     (1u << 2)
#define PFNAV_UPDATE_SET_ROTATION     (1u << 3)
#define PFNAV_UPDATE_SET_NEXT_POS     (1u << 4)
#define PFNAV_UPDATE_SET_PREV_POS     (1u << 5)
#define PFNAV_UPDATE_SET_STEP         (1u << 6)
#define PFNAV_UPDATE_SET_LEFT         (1u << 7)
#define PFNAV_UPDATE_SET_NEXT_ROT     (1u << 8)
#define PFNAV_UPDATE_SET_PREV_ROT     (1u << 9)
#define PFNAV_UPDATE_SET_DEST         (1u << 10)
#define PFNAV_UPDATE_SET_TARGET_PREV  (1u << 11)
#define PFNAV_UPDATE_SET_MOVING       (1u << 12)
#define PFNAV_UPDATE_SET_TARGET_DIR   (1u << 13)
#define PFNAV_UPDATE_TURNING_IN_PLACE (1u << 14)
/* engine_todo: what the device pass leaves to the engine for this work item */
#define PFNAV_TODO_SURROUND_QUERY     (1u << 0)  /* STATE_SURROUND_ENTITY with a live target: the branch decisions of
                                                  * movement.c:2513-2567 need the target's geometry (M_NavObjAdjacentFrom,
                                                  * M_NavClosestReachableAdjacentPosFrom: OBBs, entity tables) -- the movement
                                                  * part of the patch is complete, the state decision is the engine's */
#define PFNAV_TODO_USE_SURROUND_FIELD  (1u << 1) /* ent_update_using_surround_field (movement.c:2672): ms->using_surround_field */
#define PFNAV_TODO_DROP_SURROUND_FIELD (1u << 2) /* ... switches on / off (low / high water marks of the distance to the target) */

typedef struct pfnav_patch {            /* struct movestate_patch (movement.c:245-262); unselected fields are 0 */
    uint32_t flags;
    int32_t  next_state;                /* -1 unless SET_STATE / SET_MOVING               */
    int32_t  next_block;
    int32_t  next_attack;
    float    next_velocity[2];
    float    next_pos[3];               /* y: 0 (AIR_UNIT_HEIGHT for air units); terrain height
                                         * (M_HeightAtPoint) is render state, added by the engine  */
    float    next_rot[4];
    float    next_ppos[3];
    float    next_npos[3];
    float    next_step;
    float    next_left;
    float    next_nrot[4];
    float    next_prot[4];
    int32_t  wait_ticks_left;           /* ms->wait_ticks_left after the call (entity_compute_update counts it down in place) */
    uint32_t engine_todo;               /* PFNAV_TODO_*                                    */
    float    _padf;
    float    next_dest[2];              /* UPDATE_SET_DEST: the engine re-issues the move order (G_Move_SetDest) */
    float    next_target_prev[2];
    float    next_target_dir[4];
} pfnav_patch;                          /* 160 bytes */

/* Inspection / test entry: the two map searches of arrived() (movement.c:2170) for a flock target on `layer`, as
 * pfnav_agents_compute_updates precomputes them per (flock, layer): N_ClosestPathable (nav.c:4126) and the tile
 * centres N_IsMaximallyClose compares with (nav.c:4707). out_mc_xz: up to cap (x, z) pairs. Host code; needs
 * pfnav_route_build(layer). */
int  pfnav_route_arrival_consts(pfnav_ctx *ctx, int layer, float tx, float tz, int32_t *out_nearest_ok,
                                float *out_nearest_xz, float *out_mc_xz, size_t cap, int32_t *out_mc_n);
/* HOST array of n_agents records, uid == index. */
int  pfnav_agents_upload_movestate(pfnav_ctx *ctx, const pfnav_movestate *ms, size_t n);
/* After pfnav_agents_tick on the same stream: entity_compute_update for every work item, reading the
 * tick's new velocities and desired velocities. Needs pfnav_route_build for every layer the agents
 * use (global islands feed arrived(), movement.c:2170). Every movement state is accepted: MOVING / MOVING_IN_FORMATION
 * (formation inputs), SEEK_ENEMIES, ENTER_ENTITY_RANGE, TURNING, WAITING, ARRIVING_TO_CELL are decided on the device;
 * SURROUND_ENTITY with a live target is flagged PFNAV_TODO_SURROUND_QUERY (see there). */
int  pfnav_agents_compute_updates(pfnav_ctx *ctx, void *stream);
/* Work-item order. Blocks until the update pass has finished. */
int  pfnav_agents_read_patches(pfnav_ctx *ctx, pfnav_patch *out, size_t maxout);
/* Device-side entity_apply_update for the movestate fields (movement.c:2693-2757: state, velocity +
 * velocity history, position, interpolation fields, next rotation), then refreshes the neighbour
 * records; follow with pfnav_agents_rebuild_index before the next tick. */
int  pfnav_agents_apply_updates(pfnav_ctx *ctx, void *stream);
/* Read the (updated) entity snapshot back: agents_out / ms_out may be NULL. */
int  pfnav_agents_read_state(pfnav_ctx *ctx, pfnav_agent *agents_out, pfnav_movestate *ms_out, size_t maxout);

/* N_FlowFieldUpdate with TARGET_ENEMIES / TARGET_ENTITY straight into the pool (what N_RequestAsyncEnemySeekField /
 * N_RequestAsyncSurroundField + N_AwaitAsyncFields leave in the field cache, nav.c:3769-3960): the fields of the listed
 * chunks become the flow fields of pool destination `dest`; entities name it through pfnav_agent::aux_dest1. The
 * footprints are remembered for the in-place repairs of N_DesiredEnemySeekVelocity / N_DesiredSurroundVelocity
 * (nav.c:3647-3675, 3729-3757), which pfnav_pool_repair runs for entities standing on a tile without direction. */
int  pfnav_pool_request_entity_fields(pfnav_ctx *ctx, int dest, int layer, int ref_layer, int target_kind,
                                      const pfnav_footprint *ents, size_t nents, const int32_t *chunks_rc, size_t n,
                                      void *stream);

/* Test / tuning hook for pfnav_agents_tick: 0 = single-pass velocity kernel; 1 (default) = while LOS chains of
 * pfnav_pool_request_goals are still in flight, run the part of the update that does not depend on the
 * fields (neighbours, separation, velocity obstacles, admissible ray intersections) before joining them and
 * only the final choice after; 2 = always split. Results are identical in every mode. */
int  pfnav_set_two_phase(pfnav_ctx *ctx, int mode);

/* Test / tuning hook: cohesion pass (cohesion_force, movement.c:1653). 0 (default) = flocks of >= 20 000 members are summed
 * through the position index with an exp(-0.12 d) cut-off at 230 wu and a per-entity fall-back to the full member list
 * wherever the cut-off could matter beyond float rounding; 1 = always windowed; 2 = always the full member list. */
/* AStar_PortalGraphPath (a_star.c:429) for a batch of n searches, one device thread per search. req: 8 ints
 * per search {start chunk index, start tile r, c, end chunk index, end tile r, c, finish chunk index, finish portal index};
 * out: (4 + 3 * max_hops) ints per search {status, nhops, cost (float bits), 0, then (chunk index, portal index, local
 * island) per hop from start to finish}; status 1 = path found, 0 = none, -1 = the search outgrew its device scratch or
 * max_hops. on_device == 0 runs the host planner's own routine instead (what pfnav_route_request_path uses), for
 * comparison. Needs pfnav_route_build(layer); blocker changes must be committed. */
int  pfnav_route_graph_paths(pfnav_ctx *ctx, int layer, const int32_t *req, int n, int32_t *out, int max_hops, int on_device);
int  pfnav_set_cohesion_mode(pfnav_ctx *ctx, int mode);
/* test / tuning hook: counters of G_ClearPath_NewVelocity's retry loop (clearpath.c:702-713) since the last reset:
 * out4 = {first solves without an admissible velocity, of those with both neighbour lists non-empty, entities that
 * replayed the loop literally (order-dependent tie), solves inside those replays} */
int  pfnav_agents_clearpath_stats(pfnav_ctx *ctx, uint64_t *out4, int reset);

/* flags for pfnav_agents_tick */
#define PFNAV_TICK_VDES_FROM_POOL   (1u << 0)  /* compute vdes + has_dest_los on device (nav.c:3468, 4026) */

/* navigation_tick_task's velocity phase (movement.c:4263-4287): desired velocity gather,
 * boids steering, neighbour search, ClearPath. Asynchronous on `stream`. */
int  pfnav_agents_tick(pfnav_ctx *ctx, uint32_t flags, void *stream);

/* R_GL_MoveReadNewVelocities (render.h:672): copy the nwork new velocities (xz, 8 bytes each,
 * in work-list order) to HOST memory; blocks until the tick has finished. */
int  pfnav_agents_read_velocities(pfnav_ctx *ctx, float *out_xz, size_t maxout);
/* Diagnostics: preferred velocity (vpref), vdes and has_dest_los of each work item. */
int  pfnav_agents_read_debug(pfnav_ctx *ctx, float *out_vpref_xz, float *out_vdes_xz,
                             uint8_t *out_has_los, size_t maxout);

/* G_Pos_EntsInCircleFrom (position.c:379): neighbour query through the device spatial index,
 * returning uids in the reference's order. For parity tests of the index itself. */
int  pfnav_ents_in_circle(pfnav_ctx *ctx, float x, float z, float range, uint32_t *out,
                          int maxout, int *out_n);

/* Device-side hooks for zero-copy callers (bench `value` leg, multi-GPU all-gather):
 * raw device pointers to the resident 24-byte neighbour records (pos.xz, vel.xz, radius,
 * state << 24 | flags) in uid order, and to the velocity output (work-list order). */
int  pfnav_agents_device_ptrs(pfnav_ctx *ctx, void **d_records, void **d_velocities, size_t *n);
/* Re-build the spatial index from the (externally updated, e.g. all-gathered) record array. */
int  pfnav_agents_rebuild_index(pfnav_ctx *ctx, void *stream);

/* Optional per-kernel-group device timing (CUDA events on the launching stream) used by bench.py
 * for its roofline line. pfnav_profile_read drains the recorded intervals: ms_out / count_out have
 * 8 entries: 0 flow kernels, 1 LOS kernels, 2 spatial-index build, 3 desired velocity, 4 cohesion,
 * 5 agent velocity (ClearPath). */
int  pfnav_profile_enable(pfnav_ctx *ctx, int enable);
int  pfnav_profile_read(pfnav_ctx *ctx, float *ms_out, uint32_t *count_out);

/* Number of kernels launched by this context since creation (bench `gpu_launches`). */
uint64_t pfnav_launch_count(const pfnav_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* PFNAV_H */
