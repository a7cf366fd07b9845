// D3Q19 BGK lattice-Boltzmann sweep, f32, 24-member SoA cell (B200GEO_KERNEL_LBM_D3Q19).
//
// Replaces the per-cell update of the reference's LBM model (pull scheme + six wall states,
// src/examples/latticeboltzmann/main.cpp:62-229; SoA member set of
// src/testbed/performancetests/main.cpp:1796-1986) for the cell in oracle/models/lbm.h. The
// expression trees below are that cell's, term by term; this translation unit is compiled with
// -fmad=false and the oracle with -ffp-contract=off, so results are bit-identical.
//
// Roofline: HBM-bound. 19 populations read + 19 written = 152 B per lattice update, + 4 B of
// state read; density / velocity (16 B) are written only on the last sweep of a b200geo_step
// call unless params says otherwise (the reference cell stores them every step, 168 B/update) —
// nothing reads them on the update path, so the final grid is identical either way.
// Wall cells copy themselves and overwrite five populations; their density/velocity and every
// cell's state never change, so they are not rewritten (both buffers hold them from the load).
#include "grid.h"
#include "lbm_cell.h"

namespace b200geo {

namespace {

using namespace lbm;

#define PUT(COMP) dst[(int64_t)(COMP) * mstride + i]

// The 19 pulled populations are requested together with the state, BEFORE the wall test, so a warp
// waits for one DRAM round trip per cell instead of two (state, then populations: the first version of
// this kernel did that and reached 80 % instead of 91 % of the copy bandwidth, profiles/r1s_tuning.md).
// The loads are volatile asm (lbm_cell.h) so that neither nvcc nor ptxas sinks them back below the branch. Wall
// cells (faces only) pull values they do not use — every address is valid, the ghost ring is part of the
// array. R cells per thread along y: all 19 * R loads are in flight at once.
__device__ __forceinline__ void lbm_pull(Pulled& p, const float *__restrict__ src, int64_t i, int64_t pitch, int64_t plane,
                                         int64_t mstride)
{
    const GridHood hood = {src, i, pitch, plane, mstride};
    pull(p, hood);
}

template<bool MACRO>
__device__ __forceinline__ void lbm_collide(const Pulled& p, const float *__restrict__ src, float *__restrict__ dst, int64_t i,
                                            int64_t pitch, int64_t plane, int64_t mstride)
{
    const int s = p.state;
    float out[19];
    if (s != LIQUID) {
        // wall cell: all 24 loads are issued before the first store (one DRAM round trip for the warp, not 24
        // in a row: the x = 0 / x = max walls put one such cell into the first and last warp of EVERY row)
        const GridHood hood = {src, i, pitch, plane, mstride};
        wall(s, hood, out);
#pragma unroll
        for (int m = 0; m < 19; ++m) PUT(m) = out[m];
        return;
    }
    float rho, velX, velY, velZ;
    liquid(p, out, rho, velX, velY, velZ);
    if (MACRO) {
        PUT(DENSITY) = rho;
        PUT(VELX) = velX;
        PUT(VELY) = velY;
        PUT(VELZ) = velZ;
    }
#pragma unroll
    for (int m = 0; m < 19; ++m) PUT(m) = out[m];
}

template<bool MACRO, int R, int BX>
__global__ void __launch_bounds__(BX)
lbm_kernel_early(const float *__restrict__ src, float *__restrict__ dst, int64_t pitch, int64_t plane,
                 int64_t mstride, Box box)
{
    const int x = box.x0 + blockIdx.x * BX + threadIdx.x;
    if (x >= box.x1) return;
    const int y0 = box.y0 + blockIdx.y * R, z = box.z0 + blockIdx.z;
    const int64_t i0 = (int64_t)z * plane + (int64_t)y0 * pitch + x;
    Pulled p[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        // rows past the box (ragged last CTA) re-read the last valid row; nothing is stored for them
        const int64_t i = y0 + r < box.y1 ? i0 + r * pitch : i0;
        lbm_pull(p[r], src, i, pitch, plane, mstride);
    }
    // the states are requested LAST: the wall test cannot be scheduled above the pulls
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int64_t i = y0 + r < box.y1 ? i0 + r * pitch : i0;
        p[r].state = __float_as_int(ldg_stream(src + (int64_t)STATE * mstride + i));
    }
#pragma unroll
    for (int r = 0; r < R; ++r)
        if (y0 + r < box.y1) lbm_collide<MACRO>(p[r], src, dst, i0 + r * pitch, pitch, plane, mstride);
}


#undef PUT

}

int sweep_lbm(b200geo_grid *g, const Box& box, bool store_macroscopic, cudaStream_t s)
{
    const MemberLayout& L = g->m[0];
    int64_t mstride = g->m[1].offset / 4;
    const float *src = (const float *)g->member_ptr(0, 0) + L.origin;
    float *dst = (float *)g->member_ptr(0, 1) + L.origin;
    const int nx = box.x1 - box.x0, ny = box.y1 - box.y0, nz = box.z1 - box.z0;
    if (ny > 65535 || nz > 65535) return fail(B200GEO_ERR_OUT_OF_RANGE, "grid dimension too large");
    const int variant = g_tuning.lbm_variant;
#define LBM_LAUNCH_EARLY(R, BX)                                                                             \
    do {                                                                                                    \
        dim3 grid((nx + BX - 1) / BX, (ny + R - 1) / R, nz);                                                \
        if (store_macroscopic) lbm_kernel_early<true, R, BX><<<grid, BX, 0, s>>>(src, dst, L.pitch, L.plane, mstride, box); \
        else lbm_kernel_early<false, R, BX><<<grid, BX, 0, s>>>(src, dst, L.pitch, L.plane, mstride, box); \
    } while (0)
    if (variant == 2) {
        if (g_tuning.lbm_block == 256) LBM_LAUNCH_EARLY(2, 256); else LBM_LAUNCH_EARLY(2, 128);
    } else {
        if (g_tuning.lbm_block == 256) LBM_LAUNCH_EARLY(1, 256); else LBM_LAUNCH_EARLY(1, 128);
    }
#undef LBM_LAUNCH_EARLY
    count_launch();
    return check_cuda(cudaGetLastError(), "lbm sweep");
}

}
