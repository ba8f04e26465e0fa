/*
 * shim/gl_movement_pfnav.c -- seam B1 of SURVEY.md 8b as a COMPILED drop-in: this translation unit replaces the reference's
 * src/render/gl_movement.c (and the three position-texture entry points of gl_position.c the movement tick calls) at link
 * time. It exports the functions the movement tick pushes to the render thread in GPU mode (G_Move_SetUseGPU(true),
 * movement.c:5026), with the reference's own signatures (src/render/public/render.h:619-691), and implements them on
 * libpfnav.so: the velocities are computed by the sm_100a kernels, not by shaders/compute/movement.glsl.
 *
 * Build: against the reference's headers WHERE THEY LIE (-iquote <reference>/src), never copied: oracle/Makefile target
 * `shimb1` -> oracle/_ref/libpfnav_b1.so. tests/test_gpu_shim.py drives these entry points with buffers laid out as
 * movement.c packs them and requires the velocities pfnav_agents_tick gives for the same population.
 *
 * What the seam carries and what it does not (it is the reference's own): `struct gpu_ent_desc` has no
 * prev_pos, so ClearPath's self position is the snapshot position exactly as in movement.glsl (the CPU path uses
 * ms->prev_pos, movement.c:4351); vdes / has_dest_los arrive computed by the engine (compute_desired_velocity runs on
 * the CPU in GPU mode too). The flock buffer's member lists are cut at MAX_GPU_FLOCK_MEMBERS = 1024 (movement.c:96);
 * this shim takes membership from gpu_ent_desc::flock_id instead, so cohesion sees every member like the CPU path does.
 * cost_base / blockers come packed for all 12 layers every tick (M_NavCopyCostBasePacked, nav.c:2432); a layer is sent to
 * the device only when its bytes changed.
 */
#include "render/public/render.h"
#include "map/public/tile.h"
#include "navigation/public/nav.h"

#include "../include/pfnav.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define B1_MAX_GPU_FLOCK_MEMBERS 1024          /* movement.c:96, "must match movement.glsl" */

#ifdef PFNAV_B1_STANDALONE
/* the engine links SDL2; the stand-alone test library (oracle/Makefile `shimb1`) does not */
int SDL_AtomicSet(SDL_atomic_t *a, int v) { return __atomic_exchange_n(&a->value, v, __ATOMIC_SEQ_CST); }
#endif

/* the two shader-storage records of the seam (movement.c:341-369 == the `struct` declarations of movement.glsl) */
struct b1_flock_desc{
    uint32_t ents[B1_MAX_GPU_FLOCK_MEMBERS];
    uint32_t nmembers;
    float    target_x;
    float    target_z;
};
struct b1_ent_desc{
    vec2_t   dest;
    vec2_t   vdes;
    vec2_t   cell_pos;
    vec2_t   formation_cohesion_force;
    vec2_t   formation_align_force;
    vec2_t   formation_drag_force;
    vec2_t   pos;
    vec2_t   velocity;
    uint32_t movestate;
    uint32_t flock_id;
    uint32_t flags;
    float    speed;
    float    max_speed;
    float    radius;
    uint32_t layer;
    uint32_t has_dest_los;
    uint32_t formation_assignment_ready;
    uint32_t pad0;
};

static struct{
    pfnav_ctx          *nav;
    int                 chunk_w, chunk_h, nlayers;
    float               map_x, map_z;
    int                 hz, nwork;
    /* host copies of what R_GL_MoveUploadData was handed (its buffers live on the caller's per-frame stack) */
    uint32_t           *gpuids;     size_t ngpuids;
    struct b1_ent_desc *ents;       size_t nents;
    struct b1_flock_desc *flocks;   size_t nflocks;
    uint8_t            *cost;       size_t cost_size;
    uint16_t           *blockers;   size_t blockers_size;
    /* what the device currently holds, per layer */
    uint8_t            *dev_cost;
    uint16_t           *dev_blockers;
    bool                have_uniforms, uploaded;
}s_b1;

static void b1_die(const char *what)
{
    fprintf(stderr, "[pfnav B1 shim] %s: %s\n", what, pfnav_last_error());
    abort();                        /* gl_movement.c has no error channel either: GL errors are asserts */
}
#define B1_CHK(call) do { if((call) != PFNAV_OK) b1_die(#call); } while(0)

static void *b1_keep(void *old, const void *src, size_t bytes)
{
    void *ret = realloc(old, bytes ? bytes : 1);
    if(!ret) { fprintf(stderr, "[pfnav B1 shim] out of memory\n"); abort(); }
    memcpy(ret, src, bytes);
    return ret;
}

void R_GL_MoveUpdateUniforms(const struct map_resolution *res, vec2_t *map_pos, int *ticks_hz, int *nwork)
{
    s_b1.chunk_w = res->chunk_w;
    s_b1.chunk_h = res->chunk_h;
    s_b1.map_x = map_pos->x;
    s_b1.map_z = map_pos->z;
    s_b1.hz = *ticks_hz;
    s_b1.nwork = *nwork;
    s_b1.have_uniforms = true;
}

void R_GL_MoveUploadData(void *gpuid_buff, size_t *ndynamic_ents, void *attr_buff, size_t *attr_buffsize,
                         void *flock_buff, size_t *flock_buffsize, void *cost_base_buff, size_t *cost_base_size,
                         void *blockers_buff, size_t *blockers_size)
{
    s_b1.ngpuids = *ndynamic_ents;
    s_b1.gpuids = b1_keep(s_b1.gpuids, gpuid_buff, s_b1.ngpuids * sizeof(uint32_t));
    s_b1.nents = *attr_buffsize / sizeof(struct b1_ent_desc);
    s_b1.ents = b1_keep(s_b1.ents, attr_buff, *attr_buffsize);
    s_b1.nflocks = *flock_buffsize / sizeof(struct b1_flock_desc);
    s_b1.flocks = b1_keep(s_b1.flocks, flock_buff, *flock_buffsize);
    s_b1.cost_size = *cost_base_size;
    s_b1.cost = b1_keep(s_b1.cost, cost_base_buff, *cost_base_size);
    s_b1.blockers_size = *blockers_size;
    s_b1.blockers = b1_keep(s_b1.blockers, blockers_buff, *blockers_size);
    s_b1.uploaded = true;
}

/* the map as M_NavCopyCostBasePacked / M_NavCopyBlockersPacked lay it out: [layer][chunk_r][chunk_c][64][64] */
static void b1_sync_map(void)
{
    const size_t chunk_tiles = (size_t)s_b1.chunk_w * s_b1.chunk_h * 64 * 64;
    const int nlayers = (int)(s_b1.cost_size / chunk_tiles);
    if(!s_b1.nav || s_b1.nlayers != nlayers) {
        if(!s_b1.nav)
            B1_CHK(pfnav_create(0, &s_b1.nav));
        B1_CHK(pfnav_map_create(s_b1.nav, s_b1.chunk_w, s_b1.chunk_h, nlayers, s_b1.map_x, s_b1.map_z));
        s_b1.nlayers = nlayers;
        free(s_b1.dev_cost); free(s_b1.dev_blockers);
        s_b1.dev_cost = NULL; s_b1.dev_blockers = NULL;
    }
    B1_CHK(pfnav_map_set_pos(s_b1.nav, s_b1.map_x, s_b1.map_z));
    const bool first = !s_b1.dev_cost;
    if(first) {
        s_b1.dev_cost = malloc(s_b1.cost_size);
        s_b1.dev_blockers = malloc(s_b1.blockers_size);
        if(!s_b1.dev_cost || !s_b1.dev_blockers) { fprintf(stderr, "[pfnav B1 shim] out of memory\n"); abort(); }
    }
    for(int l = 0; l < nlayers; l++) {
        const uint8_t *c = s_b1.cost + chunk_tiles * l;
        const uint16_t *b = s_b1.blockers + chunk_tiles * l;
        if(!first && !memcmp(c, s_b1.dev_cost + chunk_tiles * l, chunk_tiles)
                  && !memcmp(b, s_b1.dev_blockers + chunk_tiles * l, chunk_tiles * sizeof(uint16_t)))
            continue;
        B1_CHK(pfnav_map_upload_layer(s_b1.nav, l, c, b, NULL));
        memcpy(s_b1.dev_cost + chunk_tiles * l, c, chunk_tiles);
        memcpy(s_b1.dev_blockers + chunk_tiles * l, b, chunk_tiles * sizeof(uint16_t));
    }
}

void R_GL_MoveDispatchWork(const size_t *nents)
{
    (void)nents;
    if(!s_b1.have_uniforms || !s_b1.uploaded) {
        fprintf(stderr, "[pfnav B1 shim] R_GL_MoveDispatchWork before R_GL_MoveUploadData / R_GL_MoveUpdateUniforms\n");
        abort();
    }
    b1_sync_map();

    const size_t n = s_b1.nents;
    pfnav_agent *agents = calloc(n ? n : 1, sizeof(pfnav_agent));
    pfnav_formation_in *form = calloc(n ? n : 1, sizeof(pfnav_formation_in));
    pfnav_flock *flocks = calloc(s_b1.nflocks ? s_b1.nflocks : 1, sizeof(pfnav_flock));
    uint32_t *work = calloc(s_b1.ngpuids ? s_b1.ngpuids : 1, sizeof(uint32_t));
    if(!agents || !form || !flocks || !work) { fprintf(stderr, "[pfnav B1 shim] out of memory\n"); abort(); }

    for(size_t f = 0; f < s_b1.nflocks; f++) {
        flocks[f].target[0] = s_b1.flocks[f].target_x;
        flocks[f].target[1] = s_b1.flocks[f].target_z;
        flocks[f].dest = -1;                 /* vdes / LOS come from the engine over this seam */
        flocks[f].layer = 0;
    }
    bool any_formation = false;
    for(size_t i = 0; i < n; i++) {          /* GPU id i + 1 == record i (movement.c:3811) */
        const struct b1_ent_desc *e = &s_b1.ents[i];
        pfnav_agent *a = &agents[i];
        a->pos[0] = e->pos.x;           a->pos[1] = e->pos.z;
        a->prev_pos[0] = e->pos.x;      a->prev_pos[1] = e->pos.z;    /* the seam has no prev_pos (see the header) */
        a->velocity[0] = e->velocity.x; a->velocity[1] = e->velocity.z;
        a->vdes[0] = e->vdes.x;         a->vdes[1] = e->vdes.z;
        a->radius = e->radius;
        a->max_speed = e->max_speed;
        a->speed = e->speed;
        a->state = e->movestate;        /* enum move_state values == enum pfnav_move_state */
        a->flags = e->flags;
        a->flock = (int32_t)e->flock_id - 1;     /* flock_id_for_ent: index + 1, 0 = none (movement.c:547) */
        a->has_dest_los = e->has_dest_los;
        if(a->flock >= 0 && (size_t)a->flock < s_b1.nflocks)
            flocks[a->flock].layer = (int32_t)e->layer;
        pfnav_formation_in *fi = &form[i];
        fi->cell_pos[0] = e->cell_pos.x;                 fi->cell_pos[1] = e->cell_pos.z;
        fi->cohesion[0] = e->formation_cohesion_force.x; fi->cohesion[1] = e->formation_cohesion_force.z;
        fi->align[0] = e->formation_align_force.x;       fi->align[1] = e->formation_align_force.z;
        fi->drag[0] = e->formation_drag_force.x;         fi->drag[1] = e->formation_drag_force.z;
        fi->target_orientation[3] = 1.0f;
        fi->cell_arrival_vdes[0] = e->vdes.x;            fi->cell_arrival_vdes[1] = e->vdes.z;   /* the record has one vdes */
        if(e->movestate == PFNAV_STATE_MOVING_IN_FORMATION || e->movestate == PFNAV_STATE_ARRIVING_TO_CELL) {
            fi->flags = PFNAV_FORM_HAS_FORMATION | (e->formation_assignment_ready ? PFNAV_FORM_ASSIGNMENT_READY : 0);
            any_formation = true;
        }
    }
    for(size_t w = 0; w < s_b1.ngpuids; w++)
        work[w] = s_b1.gpuids[w] - 1;

    B1_CHK(pfnav_agents_upload(s_b1.nav, agents, n, flocks, s_b1.nflocks, s_b1.hz));
    if(any_formation)
        B1_CHK(pfnav_agents_upload_formation(s_b1.nav, form, n));
    B1_CHK(pfnav_agents_set_work(s_b1.nav, work, s_b1.ngpuids));
    B1_CHK(pfnav_agents_tick(s_b1.nav, 0, NULL));

    free(agents); free(form); free(flocks); free(work);
}

void R_GL_MoveReadNewVelocities(void *out, const size_t *nwork, const size_t *maxout)
{
    const size_t n = *nwork < *maxout ? *nwork : *maxout;
    B1_CHK(pfnav_agents_read_velocities(s_b1.nav, (float*)out, n));     /* vec2_t[nwork] in work-list order; blocks */
}

void R_GL_MovePollCompletion(SDL_atomic_t *out)
{
    SDL_AtomicSet(out, 1);          /* pfnav_agents_read_velocities waits for the tick; there is nothing to poll for */
}

void R_GL_MoveInvalidateData(void)
{
    s_b1.uploaded = false;          /* the device buffers belong to the context and are reused by the next tick */
}

void R_GL_MoveClearState(void)
{
    if(s_b1.nav)
        pfnav_destroy(s_b1.nav);
    free(s_b1.gpuids); free(s_b1.ents); free(s_b1.flocks); free(s_b1.cost); free(s_b1.blockers);
    free(s_b1.dev_cost); free(s_b1.dev_blockers);
    memset(&s_b1, 0, sizeof(s_b1));
}

/* The position texture (gl_position.c) feeds movement.glsl's neighbour lookups; the records above carry the positions and
 * the device builds its own index, so these three are accepted and ignored. */
void R_GL_PositionsUploadData(vec3_t *posbuff, uint32_t *idbuff, const size_t *nents, const struct map *map)
{
    (void)posbuff; (void)idbuff; (void)nents; (void)map;
}

void R_GL_PositionsGetTexture(GLuint *out_tex_id)
{
    *out_tex_id = 0;
}

void R_GL_PositionsInvalidateData(void)
{
}
