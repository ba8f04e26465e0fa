// pfnav_mgpu.cu -- multi-GPU: entities (and with them their flocks' destinations) are partitioned by contiguous
// index range, the reference's own fork-join split (src/game/movement.c:3751-3762: equal ranges of the work
// array, one per task); every GPU updates its own range and needs every potential neighbour's 24-byte record
// {pos, vel, radius, state|flags} -> ONE all-gather of those records per tick, nothing else.
//
// Two transports behind the same per-tick call:
//   * NCCL, one process per GPU (pfnav_mgpu_init): libnccl.so.2 is resolved at run time with dlopen (the copy the
//     host program already loaded, e.g. torch's, else the system one); ncclAllGather in place on the record array,
//     stream-ordered on the caller's stream, NVLink / NVSwitch underneath.
//   * in-process group (pfnav_group_create): one engine process driving several contexts (on one or several
//     devices) from one thread, which is how the engine itself would use 8 GPUs; the ranges are pulled with
//     cudaMemcpyPeerAsync (NVLink peer copies between devices), ordered by events.
// The map, the field pool of the rank's own destinations and the spatial index are per rank; the index is rebuilt
// from the gathered records on every rank (identical order on all of them).
#include "pfnav_internal.cuh"
#include <dlfcn.h>
#include <string.h>
#include <algorithm>

// ---- minimal NCCL surface (nccl.h:  ncclUniqueId 128 bytes, ncclResult_t, ncclDataType_t ncclInt8 = 0) ----
typedef struct { char internal[128]; } pf_nccl_id;
typedef void *pf_nccl_comm;
typedef int (*fn_ncclGetUniqueId)(pf_nccl_id *);
typedef int (*fn_ncclCommInitRank)(pf_nccl_comm *, int, pf_nccl_id, int);
typedef int (*fn_ncclAllGather)(const void *, void *, size_t, int, pf_nccl_comm, cudaStream_t);
typedef int (*fn_ncclBroadcast)(const void *, void *, size_t, int, int, pf_nccl_comm, cudaStream_t);
typedef int (*fn_ncclGroup)(void);
typedef int (*fn_ncclCommDestroy)(pf_nccl_comm);
typedef const char *(*fn_ncclGetErrorString)(int);

struct pf_nccl_api {
    void *lib = nullptr;
    fn_ncclGetUniqueId GetUniqueId = nullptr; fn_ncclCommInitRank CommInitRank = nullptr;
    fn_ncclAllGather AllGather = nullptr; fn_ncclBroadcast Broadcast = nullptr;
    fn_ncclGroup GroupStart = nullptr, GroupEnd = nullptr;
    fn_ncclCommDestroy CommDestroy = nullptr; fn_ncclGetErrorString GetErrorString = nullptr;
};

static int nccl_load(pf_nccl_api &api)
{
    if (api.lib) return 0;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) { api.lib = dlopen(nm, RTLD_NOW | RTLD_NOLOAD); if (api.lib) break; }     // already in the process?
    for (const char *nm : names) { if (api.lib) break; api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); }
    if (!api.lib) { pfnav_set_error("pfnav_mgpu: libnccl.so.2 not found (%s)", dlerror()); return PFNAV_ERR_STATE; }
#define PF_SYM(name) api.name = (fn_nccl##name)dlsym(api.lib, "nccl" #name)
    PF_SYM(GetUniqueId); PF_SYM(CommInitRank); PF_SYM(AllGather); PF_SYM(Broadcast); PF_SYM(CommDestroy); PF_SYM(GetErrorString);
#undef PF_SYM
    api.GroupStart = (fn_ncclGroup)dlsym(api.lib, "ncclGroupStart");
    api.GroupEnd = (fn_ncclGroup)dlsym(api.lib, "ncclGroupEnd");
    if (!api.GetUniqueId || !api.CommInitRank || !api.AllGather || !api.Broadcast || !api.GroupStart || !api.GroupEnd || !api.CommDestroy) {
        pfnav_set_error("pfnav_mgpu: libnccl.so.2 lacks a required symbol");
        return PFNAV_ERR_STATE;
    }
    return 0;
}
static pf_nccl_api g_nccl;        // resolved symbols only (immutable after the first load)

struct pfnav_group {
    std::vector<pfnav_ctx *> ctxs;
    std::vector<cudaEvent_t> ready, pulled;      // per member: own range final / every peer range copied
};

struct pf_mgpu {
    int rank = 0, world = 1;
    pf_nccl_comm comm = nullptr;                 // NCCL transport
    pfnav_group *group = nullptr;                // in-process transport
};

#define PF_NCCL(call)                                                                                  \
    do {                                                                                               \
        int _r = (call);                                                                               \
        if (_r != 0) {                                                                                 \
            pfnav_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call,                              \
                            g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "nccl error");         \
            return PFNAV_ERR_CUDA;                                                                     \
        }                                                                                              \
    } while (0)

// balanced contiguous ranges (movement.c:3751-3762 splits the work array the same way)
extern "C" int pfnav_mgpu_shard_range(size_t n_total, int rank, int world, size_t *lo, size_t *hi)
{
    PF_ARG(world > 0 && rank >= 0 && rank < world && lo && hi, "rank / world");
    const size_t base = n_total / world, rem = n_total % world;
    *lo = (size_t)rank * base + std::min<size_t>(rank, rem);
    *hi = *lo + base + ((size_t)rank < rem ? 1 : 0);
    return PFNAV_OK;
}

extern "C" int pfnav_mgpu_unique_id(void *out_id)
{
    PF_ARG(out_id, "out_id");
    int rc = nccl_load(g_nccl);
    if (rc) return rc;
    pf_nccl_id id;
    PF_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(out_id, &id, sizeof(id));
    return PFNAV_OK;
}

extern "C" int pfnav_mgpu_init(pfnav_ctx *ctx, int rank, int world, const void *id)
{
    PF_ARG(ctx && id && world >= 1 && rank >= 0 && rank < world, "args");
    PF_NEED_DEVICE(ctx);
    PF_ARG(!ctx->mgpu, "context already belongs to a multi-GPU job");
    int rc = nccl_load(g_nccl);
    if (rc) return rc;
    PF_CUDA(cudaSetDevice(ctx->device));
    pf_mgpu *m = new pf_mgpu();
    m->rank = rank; m->world = world;
    pf_nccl_id nid;
    memcpy(&nid, id, sizeof(nid));
    int r = g_nccl.CommInitRank(&m->comm, world, nid, rank);
    if (r != 0) {
        pfnav_set_error("pfnav_mgpu_init: ncclCommInitRank -> %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error");
        delete m;
        return PFNAV_ERR_CUDA;
    }
    ctx->mgpu = m;
    return PFNAV_OK;
}

extern "C" int pfnav_mgpu_finalize(pfnav_ctx *ctx)
{
    PF_ARG(ctx, "ctx");
    pf_mgpu *m = (pf_mgpu *)ctx->mgpu;
    if (!m) return PFNAV_OK;
    PF_ARG(!m->group, "member of an in-process group: pfnav_group_destroy");
    if (ctx->device >= 0) { cudaSetDevice(ctx->device); cudaDeviceSynchronize(); }
    if (m->comm) g_nccl.CommDestroy(m->comm);
    delete m;
    ctx->mgpu = nullptr;
    return PFNAV_OK;
}

// all-gather of one per-entity column (elem_bytes per entity, uid order) in place: every rank holds its own range
// and receives the others'. Ranges are the balanced split, so they differ by at most one entity: equal ranges use
// ncclAllGather, unequal ones a group of ncclBroadcast (the all-gather-v idiom).
static int nccl_allgather(pfnav_ctx *ctx, pf_mgpu *m, void *d_buf, size_t elem_bytes, cudaStream_t st)
{
    const size_t n = ctx->n_agents;
    if (n % m->world == 0) {
        const size_t cnt = n / m->world * elem_bytes;
        PF_NCCL(g_nccl.AllGather((const uint8_t *)d_buf + (size_t)m->rank * cnt, d_buf, cnt, /*ncclInt8*/ 0, m->comm, st));
        return 0;
    }
    PF_NCCL(g_nccl.GroupStart());
    for (int r = 0; r < m->world; r++) {
        size_t lo, hi;
        pfnav_mgpu_shard_range(n, r, m->world, &lo, &hi);
        uint8_t *p = (uint8_t *)d_buf + lo * elem_bytes;
        PF_NCCL(g_nccl.Broadcast(p, p, (hi - lo) * elem_bytes, 0, r, m->comm, st));
    }
    PF_NCCL(g_nccl.GroupEnd());
    return 0;
}

int pfnav_mgpu_allgather(pfnav_ctx *ctx, void *d_buf, size_t elem_bytes, cudaStream_t st)
{
    pf_mgpu *m = (pf_mgpu *)ctx->mgpu;
    if (!m || m->world == 1) return 0;
    PF_ARG(m->comm, "in-process groups gather through pfnav_group_gather");
    return nccl_allgather(ctx, m, d_buf, elem_bytes, st);
}

static int check_shard(pfnav_ctx *ctx, const pf_mgpu *m)
{
    size_t lo, hi;
    pfnav_mgpu_shard_range(ctx->n_agents, m->rank, m->world, &lo, &hi);
    PF_ARG(ctx->d_records && lo == ctx->shard_lo && hi == ctx->shard_hi,
           "upload this rank's own range first (pfnav_agents_upload_shard with the range of pfnav_mgpu_shard_range)");
    return 0;
}

// The per-tick collective (NCCL transport): all-gather of the 24-byte records -- and of the flock id column when
// membership may have changed since the last gather -- then the whole-population part of the snapshot on this rank
// (member lists if stale, spatial index). Asynchronous on `stream` unless member lists are rebuilt.
extern "C" int pfnav_mgpu_gather(pfnav_ctx *ctx, void *stream)
{
    PF_ARG(ctx && ctx->mgpu, "pfnav_mgpu_init not called");
    PF_NEED_DEVICE(ctx);
    pf_mgpu *m = (pf_mgpu *)ctx->mgpu;
    PF_ARG(!m->group, "member of an in-process group: pfnav_group_gather");
    int rc = check_shard(ctx, m);
    if (rc) return rc;
    PF_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = pf_stream(ctx, stream);
    if ((rc = nccl_allgather(ctx, m, ctx->d_records, sizeof(pf_record), st))) return rc;
    const bool members = ctx->members_stale;
    if (members && (rc = nccl_allgather(ctx, m, ctx->d_flock_of, sizeof(int32_t), st))) return rc;
    ctx->members_stale = false;
    if (members) return pfnav_agents_finish_snapshot(ctx, st, true);
    return pfnav_agents_rebuild_index(ctx, st);
}

// ------------------------------------------------------------------------------------------
// In-process group: several contexts of ONE process (one per GPU, or several on one GPU for tests), driven from
// one thread. Member i is rank i of world = n.
// ------------------------------------------------------------------------------------------
extern "C" int pfnav_group_create(pfnav_ctx **ctxs, int world, pfnav_group **out)
{
    PF_ARG(ctxs && out && world >= 1, "args");
    for (int i = 0; i < world; i++) {
        PF_ARG(ctxs[i] && !ctxs[i]->mgpu, "context missing or already in a multi-GPU job");
        PF_NEED_DEVICE(ctxs[i]);
    }
    pfnav_group *g = new pfnav_group();
    g->ctxs.assign(ctxs, ctxs + world);
    g->ready.resize(world); g->pulled.resize(world);
    for (int i = 0; i < world; i++) {
        PF_CUDA(cudaSetDevice(ctxs[i]->device));
        PF_CUDA(cudaEventCreateWithFlags(&g->ready[i], cudaEventDisableTiming));
        PF_CUDA(cudaEventCreateWithFlags(&g->pulled[i], cudaEventDisableTiming));
        for (int j = 0; j < world; j++) {                 // NVLink peer copies between the members' devices
            if (ctxs[j]->device == ctxs[i]->device) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, ctxs[i]->device, ctxs[j]->device);
            if (can && cudaDeviceEnablePeerAccess(ctxs[j]->device, 0) != cudaSuccess) cudaGetLastError();   // already enabled
        }
        pf_mgpu *m = new pf_mgpu();
        m->rank = i; m->world = world; m->group = g;
        ctxs[i]->mgpu = m;
    }
    *out = g;
    return PFNAV_OK;
}

extern "C" void pfnav_group_destroy(pfnav_group *g)
{
    if (!g) return;
    for (size_t i = 0; i < g->ctxs.size(); i++) {
        pfnav_ctx *c = g->ctxs[i];
        cudaSetDevice(c->device);
        cudaDeviceSynchronize();
        cudaEventDestroy(g->ready[i]); cudaEventDestroy(g->pulled[i]);
        delete (pf_mgpu *)c->mgpu;
        c->mgpu = nullptr;
    }
    delete g;
}

// The per-tick collective of the group: every member pulls the other members' ranges of the record array (and of
// the flock id column when membership may have changed) and finishes its snapshot. All on the members' own streams.
extern "C" int pfnav_group_gather(pfnav_group *g)
{
    PF_ARG(g, "group");
    const int world = (int)g->ctxs.size();
    int rc;
    bool members = false;
    for (int i = 0; i < world; i++) {
        pfnav_ctx *c = g->ctxs[i];
        PF_ARG(c->n_agents == g->ctxs[0]->n_agents, "members hold populations of different sizes");
        if ((rc = check_shard(c, (pf_mgpu *)c->mgpu))) return rc;
        members |= c->members_stale;
        PF_CUDA(cudaSetDevice(c->device));
        PF_CUDA(cudaEventRecord(g->ready[i], c->tick_stream));       // my range is final once my stream gets here
    }
    for (int i = 0; i < world; i++) {
        pfnav_ctx *c = g->ctxs[i];
        PF_CUDA(cudaSetDevice(c->device));
        for (int p = 0; p < world; p++) {
            if (p == i) continue;
            pfnav_ctx *src = g->ctxs[p];
            PF_CUDA(cudaStreamWaitEvent(c->tick_stream, g->ready[p], 0));
            const size_t lo = src->shard_lo, cnt = src->shard_hi - src->shard_lo;
            if (!cnt) continue;
            PF_CUDA(cudaMemcpyPeerAsync(c->d_records + lo, c->device, src->d_records + lo, src->device, cnt * sizeof(pf_record), c->tick_stream));
            if (members)
                PF_CUDA(cudaMemcpyPeerAsync(c->d_flock_of + lo, c->device, src->d_flock_of + lo, src->device, cnt * 4, c->tick_stream));
        }
        PF_CUDA(cudaEventRecord(g->pulled[i], c->tick_stream));
    }
    for (int i = 0; i < world; i++) {
        pfnav_ctx *c = g->ctxs[i];
        PF_CUDA(cudaSetDevice(c->device));
        // whatever member i does next to its own range (the state apply of the next tick) must wait until every
        // peer has copied the current version
        for (int p = 0; p < world; p++)
            if (p != i) PF_CUDA(cudaStreamWaitEvent(c->tick_stream, g->pulled[p], 0));
        c->members_stale = false;
        rc = members ? pfnav_agents_finish_snapshot(c, c->tick_stream, true) : pfnav_agents_rebuild_index(c, c->tick_stream);
        if (rc) return rc;
    }
    return PFNAV_OK;
}
