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

namespace b200geo {

namespace {

enum { C, N, E, W, S, T, B, NW, SW, NE, SE, TW, BW, TE, BE, TN, BN, TS, BS, DENSITY, VELX, VELY, VELZ, STATE };
enum { LIQUID, WEST_NOSLIP, EAST_NOSLIP, TOP, BOTTOM, NORTH_ACC, SOUTH_NOSLIP };

#define GET_COMP(X, Y, Z, COMP) src[(int64_t)(COMP) * mstride + i + (X) + (Y) * pitch + (Z) * plane]
#define PUT(COMP) dst[(int64_t)(COMP) * mstride + i]
#define SQR(X) ((X) * (X))

// The 19 pulled populations are requested together with the state, BEFORE the wall test, so a warp
// waits for one DRAM round trip per cell instead of two (state, then populations: the first version of
// this kernel did that and reached 80 % instead of 91 % of the copy bandwidth, profiles/r1s_tuning.md).
// The loads are volatile asm so that neither nvcc nor ptxas sinks them back below the branch. Wall cells (faces only) pull values they do not use — every address is valid,
// the ghost ring is part of the array. R cells per thread along y: all 19 * R loads are in flight
// at once.
__device__ __forceinline__ float ldg_stream(const float *p)
{
    float v;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

struct Pulled {
    float fC, fN, fS, fE, fW, fT, fB, fNW, fSW, fNE, fSE, fTW, fBW, fTE, fBE, fTN, fBN, fTS, fBS;
    int state;
};

__device__ __forceinline__ void lbm_pull(Pulled& p, const float *__restrict__ src, int64_t i, int64_t pitch, int64_t plane,
                                         int64_t mstride)
{
#define PULL(X, Y, Z, COMP) ldg_stream(src + (int64_t)(COMP) * mstride + i + (X) + (Y) * pitch + (Z) * plane)
    p.fC  = PULL( 0, 0, 0, C);
    p.fN  = PULL( 0,-1, 0, N);   p.fS  = PULL( 0, 1, 0, S);
    p.fE  = PULL(-1, 0, 0, E);   p.fW  = PULL( 1, 0, 0, W);
    p.fT  = PULL( 0, 0,-1, T);   p.fB  = PULL( 0, 0, 1, B);
    p.fNW = PULL( 1,-1, 0, NW);  p.fSW = PULL( 1, 1, 0, SW);
    p.fNE = PULL(-1,-1, 0, NE);  p.fSE = PULL(-1, 1, 0, SE);
    p.fTW = PULL( 1, 0,-1, TW);  p.fBW = PULL( 1, 0, 1, BW);
    p.fTE = PULL(-1, 0,-1, TE);  p.fBE = PULL(-1, 0, 1, BE);
    p.fTN = PULL( 0,-1,-1, TN);  p.fBN = PULL( 0,-1, 1, BN);
    p.fTS = PULL( 0, 1,-1, TS);  p.fBS = PULL( 0, 1, 1, BS);
#undef PULL
}

template<bool MACRO>
__device__ __forceinline__ void lbm_collide(const Pulled& p, const float *__restrict__ src, float *__restrict__ dst, int64_t i,
                                            int64_t pitch, int64_t plane, int64_t mstride)
{
    const int s = p.state;
    if (s != LIQUID) {
        // wall cell: copy itself, overwrite the five populations that point into the fluid. All 24
        // loads are issued before the first store (one DRAM round trip for the warp, not 24 in a row:
        // the x = 0 / x = max walls put one such cell into the first and last warp of EVERY row)
#define OWN(X, Y, Z, COMP) ldg_stream(src + (int64_t)(COMP) * mstride + i + (X) + (Y) * pitch + (Z) * plane)
        float own[19];
#pragma unroll
        for (int m = 0; m < 19; ++m) own[m] = OWN(0, 0, 0, m);
        switch (s) {
        case WEST_NOSLIP:
            own[E]  = OWN(1, 0,  0, W);
            own[NE] = OWN(1, 1,  0, SW);
            own[SE] = OWN(1,-1,  0, NW);
            own[TE] = OWN(1, 0,  1, BW);
            own[BE] = OWN(1, 0, -1, TW);
            break;
        case EAST_NOSLIP:
            own[W]  = OWN(-1, 0, 0, E);
            own[NW] = OWN(-1, 0, 1, SE);
            own[SW] = OWN(-1,-1, 0, NE);
            own[TW] = OWN(-1, 0, 1, BE);
            own[BW] = OWN(-1, 0,-1, TE);
            break;
        case TOP:
            own[B]  = OWN(0, 0,-1, T);
            own[BE] = OWN(1, 0,-1, TW);
            own[BW] = OWN(-1,0,-1, TE);
            own[BN] = OWN(0, 1,-1, TS);
            own[BS] = OWN(0,-1,-1, TN);
            break;
        case BOTTOM:
            own[T]  = OWN(0, 0, 1, B);
            own[TE] = OWN(1, 0, 1, BW);
            own[TW] = OWN(-1,0, 1, BE);
            own[TN] = OWN(0, 1, 1, BS);
            own[TS] = OWN(0,-1, 1, BN);
            break;
        case NORTH_ACC: {
            const float w_1 = 0.01f;
            own[S]  = OWN(0,-1, 0, N);
            own[SE] = OWN(1,-1, 0, NW) + 6.0f * w_1 * 0.1f;
            own[SW] = OWN(-1,-1,0, NE) - 6.0f * w_1 * 0.1f;
            own[TS] = OWN(0,-1, 1, BN);
            own[BS] = OWN(0,-1,-1, TN);
            break;
        }
        case SOUTH_NOSLIP:
            own[N]  = OWN(0, 1, 0, S);
            own[NE] = OWN(1, 1, 0, SW);
            own[NW] = OWN(-1,1, 0, SE);
            own[TN] = OWN(0, 1, 1, BS);
            own[BN] = OWN(0, 1,-1, TS);
            break;
        }
#undef OWN
#pragma unroll
        for (int m = 0; m < 19; ++m) PUT(m) = own[m];
        return;
    }

    const float omega     = (float)(1.0 / 1.7);
    const float omega_trm = 1.0f - omega;
    const float omega_w0  = (float)(3.0 * 1.0 / 3.0)  * omega;
    const float omega_w1  = (float)(3.0 * 1.0 / 18.0) * omega;
    const float omega_w2  = (float)(3.0 * 1.0 / 36.0) * omega;
    const float one_third = (float)(1.0 / 3.0);

    const float fC = p.fC, fN = p.fN, fS = p.fS, fE = p.fE, fW = p.fW, fT = p.fT, fB = p.fB;
    const float fNW = p.fNW, fSW = p.fSW, fNE = p.fNE, fSE = p.fSE, fTW = p.fTW, fBW = p.fBW, fTE = p.fTE, fBE = p.fBE;
    const float fTN = p.fTN, fBN = p.fBN, fTS = p.fTS, fBS = p.fBS;

    float velX, velY, velZ;
    velX = fE + fNE + fSE + fTE + fBE;
    velY = fN + fNW + fTN + fBN;
    velZ = fT + fTS + fTW;

    const float rho = fC + fS + fW + fB + fSW + fBS + fBW + velX + velY + velZ;
    velX = velX - fW - fNW - fSW - fTW - fBW;
    velY = velY + fNE - fS - fSW - fSE - fTS - fBS;
    velZ = velZ + fTN + fTE - fB - fBN - fBS - fBW - fBE;

    if (MACRO) {
        PUT(DENSITY) = rho;
        PUT(VELX) = velX;
        PUT(VELY) = velY;
        PUT(VELZ) = velZ;
    }

    const float dir_indep_trm = one_third * rho - 0.5f * (velX * velX + velY * velY + velZ * velZ);

    PUT(C)  = omega_trm * fC + omega_w0 * (dir_indep_trm);

    PUT(NW) = omega_trm * fNW + omega_w2 * (dir_indep_trm - (velX - velY) + 1.5f * SQR(velX - velY));
    PUT(SE) = omega_trm * fSE + omega_w2 * (dir_indep_trm + (velX - velY) + 1.5f * SQR(velX - velY));
    PUT(NE) = omega_trm * fNE + omega_w2 * (dir_indep_trm + (velX + velY) + 1.5f * SQR(velX + velY));
    PUT(SW) = omega_trm * fSW + omega_w2 * (dir_indep_trm - (velX + velY) + 1.5f * SQR(velX + velY));

    PUT(TW) = omega_trm * fTW + omega_w2 * (dir_indep_trm - (velX - velZ) + 1.5f * SQR(velX - velZ));
    PUT(BE) = omega_trm * fBE + omega_w2 * (dir_indep_trm + (velX - velZ) + 1.5f * SQR(velX - velZ));
    PUT(TE) = omega_trm * fTE + omega_w2 * (dir_indep_trm + (velX + velZ) + 1.5f * SQR(velX + velZ));
    PUT(BW) = omega_trm * fBW + omega_w2 * (dir_indep_trm - (velX + velZ) + 1.5f * SQR(velX + velZ));

    PUT(TS) = omega_trm * fTS + omega_w2 * (dir_indep_trm - (velY - velZ) + 1.5f * SQR(velY - velZ));
    PUT(BN) = omega_trm * fBN + omega_w2 * (dir_indep_trm + (velY - velZ) + 1.5f * SQR(velY - velZ));
    PUT(TN) = omega_trm * fTN + omega_w2 * (dir_indep_trm + (velY + velZ) + 1.5f * SQR(velY + velZ));
    PUT(BS) = omega_trm * fBS + omega_w2 * (dir_indep_trm - (velY + velZ) + 1.5f * SQR(velY + velZ));

    PUT(N) = omega_trm * fN + omega_w1 * (dir_indep_trm + velY + 1.5f * SQR(velY));
    PUT(S) = omega_trm * fS + omega_w1 * (dir_indep_trm - velY + 1.5f * SQR(velY));
    PUT(E) = omega_trm * fE + omega_w1 * (dir_indep_trm + velX + 1.5f * SQR(velX));
    PUT(W) = omega_trm * fW + omega_w1 * (dir_indep_trm - velX + 1.5f * SQR(velX));
    PUT(T) = omega_trm * fT + omega_w1 * (dir_indep_trm + velZ + 1.5f * SQR(velZ));
    PUT(B) = omega_trm * fB + omega_w1 * (dir_indep_trm - velZ + 1.5f * SQR(velZ));
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


#undef GET_COMP
#undef PUT
#undef SQR

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
