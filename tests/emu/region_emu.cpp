// tests/emu/region_emu.cpp -- TEST INFRASTRUCTURE. Runs the device source of k_region_fields
// (permafrost-engine_b200/csrc/pfnav_region_kernel.cuh, unmodified) as ONE thread block on CPU threads, so that the
// kernel's index arithmetic and fixed-point logic are checked against the golden vectors in the CPU test suite too.
// The CUDA keywords the header uses are defined as plain C++; __syncthreads is a pthread barrier. This is not a
// product path: nothing in libpfnav.so links or calls it, and it is built only by tests/test_region_emu.py.
#include <pthread.h>
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>
#include "../../include/pfnav.h"

#define RG_THREADS 32
#define __device__
#define __global__
#define __forceinline__ inline
#define __restrict__
#define __shared__
#define __launch_bounds__(...)

struct emu_dim3 { unsigned x, y, z; };
static thread_local emu_dim3 threadIdx;
static const emu_dim3 blockIdx = {0, 0, 0}, gridDim = {1, 1, 1};
static pthread_barrier_t g_barrier;
static std::atomic<int> g_or{0};
alignas(16) uint32_t rg_smem[PFNAV_REGION_DIM_MAX * PFNAV_REGION_DIM_MAX * 6 / 4];

static inline void __syncthreads() { pthread_barrier_wait(&g_barrier); }
static inline int __syncthreads_or(int p)
{
    if (p) g_or.store(1);
    pthread_barrier_wait(&g_barrier);
    const int r = g_or.load();
    pthread_barrier_wait(&g_barrier);
    if (threadIdx.x == 0) g_or.store(0);
    pthread_barrier_wait(&g_barrier);
    return r;
}
using std::min;
using std::max;

#include "../../permafrost-engine_b200/csrc/pfnav_region_kernel.cuh"

// cost / blk / fmask: row-major layer images [layer][H64][W64], exactly what the device holds
extern "C" int emu_region_fields(const uint8_t *cost, const uint16_t *blk, const uint16_t *fmask, int W64, int H64, int dim,
                                 const pfnav_region_req *reqs, int n, const int32_t *seeds, const int32_t *overlay, uint8_t *fields,
                                 int chunk_out)
{
    RegionGrids g = { cost, blk, fmask, W64, H64 };
    pthread_barrier_init(&g_barrier, nullptr, RG_THREADS);
    std::vector<std::thread> th;
    for (unsigned t = 0; t < RG_THREADS; t++)
        th.emplace_back([=]() {
            threadIdx = {t, 0, 0};
            k_region_fields(g, dim, reqs, n, seeds, overlay, fields, chunk_out);
        });
    for (auto &t : th) t.join();
    pthread_barrier_destroy(&g_barrier);
    return 0;
}
