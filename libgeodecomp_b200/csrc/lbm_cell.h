// The D3Q19 BGK cell of oracle/models/lbm.h as device code, shared by the one-sweep kernel (lbm.cu) and the
// kernel that fuses two sweeps (lbm_tb.cu).
//
// Follows the reference's LBM model term by term (pull scheme + six wall states,
// src/examples/latticeboltzmann/main.cpp:62-229; SoA member set of
// src/testbed/performancetests/main.cpp:1796-1986). Where a neighbour's population comes from is left to an
// accessor `a.template get<X, Y, Z, COMP>()` (global memory for a sweep that reads the grid, shared memory for
// the second of two fused sweeps); the expression trees are the same in both, and both translation units are
// compiled with -fmad=false, so every sweep is bit-identical to the -ffp-contract=off oracle.
#ifndef B200GEO_CSRC_LBM_CELL_H
#define B200GEO_CSRC_LBM_CELL_H

namespace b200geo {
namespace lbm {

enum { C, N, E, W, S, T, B, NW, SW, NE, SE, TW, BW, TE, BE, TN, BN, TS, BS, DENSITY, VELX, VELY, VELZ, STATE };
enum { LIQUID, WEST_NOSLIP, EAST_NOSLIP, TOP, BOTTOM, NORTH_ACC, SOUTH_NOSLIP };

struct Pulled {
    float fC, fN, fS, fE, fW, fT, fB, fNW, fSW, fNE, fSE, fTW, fBW, fTE, fBE, fTN, fBN, fTS, fBS;
    int state;
};

// the 19 populations that stream into a cell (src/examples/latticeboltzmann/main.cpp:101-119: GET_COMP offsets)
template<class A>
__device__ __forceinline__ void pull(Pulled& p, const A& a)
{
    p.fC  = a.template get< 0, 0, 0, C>();
    p.fN  = a.template get< 0,-1, 0, N>();   p.fS  = a.template get< 0, 1, 0, S>();
    p.fE  = a.template get<-1, 0, 0, E>();   p.fW  = a.template get< 1, 0, 0, W>();
    p.fT  = a.template get< 0, 0,-1, T>();   p.fB  = a.template get< 0, 0, 1, B>();
    p.fNW = a.template get< 1,-1, 0, NW>();  p.fSW = a.template get< 1, 1, 0, SW>();
    p.fNE = a.template get<-1,-1, 0, NE>();  p.fSE = a.template get<-1, 1, 0, SE>();
    p.fTW = a.template get< 1, 0,-1, TW>();  p.fBW = a.template get< 1, 0, 1, BW>();
    p.fTE = a.template get<-1, 0,-1, TE>();  p.fBE = a.template get<-1, 0, 1, BE>();
    p.fTN = a.template get< 0,-1,-1, TN>();  p.fBN = a.template get< 0,-1, 1, BN>();
    p.fTS = a.template get< 0, 1,-1, TS>();  p.fBS = a.template get< 0, 1, 1, BS>();
}

// the cell's own 19 populations (wall cells and cells outside the simulation area copy themselves)
template<class A>
__device__ __forceinline__ void own_populations(const A& a, float (&own)[19])
{
    own[C]  = a.template get<0, 0, 0, C>();   own[N]  = a.template get<0, 0, 0, N>();
    own[E]  = a.template get<0, 0, 0, E>();   own[W]  = a.template get<0, 0, 0, W>();
    own[S]  = a.template get<0, 0, 0, S>();   own[T]  = a.template get<0, 0, 0, T>();
    own[B]  = a.template get<0, 0, 0, B>();   own[NW] = a.template get<0, 0, 0, NW>();
    own[SW] = a.template get<0, 0, 0, SW>();  own[NE] = a.template get<0, 0, 0, NE>();
    own[SE] = a.template get<0, 0, 0, SE>();  own[TW] = a.template get<0, 0, 0, TW>();
    own[BW] = a.template get<0, 0, 0, BW>();  own[TE] = a.template get<0, 0, 0, TE>();
    own[BE] = a.template get<0, 0, 0, BE>();  own[TN] = a.template get<0, 0, 0, TN>();
    own[BN] = a.template get<0, 0, 0, BN>();  own[TS] = a.template get<0, 0, 0, TS>();
    own[BS] = a.template get<0, 0, 0, BS>();
}

// wall cell: copy itself, overwrite the five populations that point into the fluid
// (src/examples/latticeboltzmann/main.cpp:160-229, including the (-1, 0, 1) offset of EAST_NOSLIP's NW entry)
template<class A>
__device__ __forceinline__ void wall(int s, const A& a, float (&own)[19])
{
    own_populations(a, own);
    switch (s) {
    case WEST_NOSLIP:
        own[E]  = a.template get<1, 0,  0, W>();
        own[NE] = a.template get<1, 1,  0, SW>();
        own[SE] = a.template get<1,-1,  0, NW>();
        own[TE] = a.template get<1, 0,  1, BW>();
        own[BE] = a.template get<1, 0, -1, TW>();
        break;
    case EAST_NOSLIP:
        own[W]  = a.template get<-1, 0, 0, E>();
        own[NW] = a.template get<-1, 0, 1, SE>();
        own[SW] = a.template get<-1,-1, 0, NE>();
        own[TW] = a.template get<-1, 0, 1, BE>();
        own[BW] = a.template get<-1, 0,-1, TE>();
        break;
    case TOP:
        own[B]  = a.template get<0, 0,-1, T>();
        own[BE] = a.template get<1, 0,-1, TW>();
        own[BW] = a.template get<-1,0,-1, TE>();
        own[BN] = a.template get<0, 1,-1, TS>();
        own[BS] = a.template get<0,-1,-1, TN>();
        break;
    case BOTTOM:
        own[T]  = a.template get<0, 0, 1, B>();
        own[TE] = a.template get<1, 0, 1, BW>();
        own[TW] = a.template get<-1,0, 1, BE>();
        own[TN] = a.template get<0, 1, 1, BS>();
        own[TS] = a.template get<0,-1, 1, BN>();
        break;
    case NORTH_ACC: {
        const float w_1 = 0.01f;
        own[S]  = a.template get<0,-1, 0, N>();
        own[SE] = a.template get<1,-1, 0, NW>() + 6.0f * w_1 * 0.1f;
        own[SW] = a.template get<-1,-1,0, NE>() - 6.0f * w_1 * 0.1f;
        own[TS] = a.template get<0,-1, 1, BN>();
        own[BS] = a.template get<0,-1,-1, TN>();
        break;
    }
    case SOUTH_NOSLIP:
        own[N]  = a.template get<0, 1, 0, S>();
        own[NE] = a.template get<1, 1, 0, SW>();
        own[NW] = a.template get<-1,1, 0, SE>();
        own[TN] = a.template get<0, 1, 1, BS>();
        own[BN] = a.template get<0, 1,-1, TS>();
        break;
    }
}

#define B200GEO_LBM_SQR(X) ((X) * (X))

// liquid cell: BGK collision of the pulled populations (src/examples/latticeboltzmann/main.cpp:95-158)
__device__ __forceinline__ void liquid(const Pulled& p, float (&out)[19], float& rho, float& velX, float& velY, float& velZ)
{
    const float omega     = (float)(1.0 / 1.7);
    const float omega_trm = 1.0f - omega;
    const float omega_w0  = (float)(3.0 * 1.0 / 3.0)  * omega;
    const float omega_w1  = (float)(3.0 * 1.0 / 18.0) * omega;
    const float omega_w2  = (float)(3.0 * 1.0 / 36.0) * omega;
    const float one_third = (float)(1.0 / 3.0);

    const float fC = p.fC, fN = p.fN, fS = p.fS, fE = p.fE, fW = p.fW, fT = p.fT, fB = p.fB;
    const float fNW = p.fNW, fSW = p.fSW, fNE = p.fNE, fSE = p.fSE, fTW = p.fTW, fBW = p.fBW, fTE = p.fTE, fBE = p.fBE;
    const float fTN = p.fTN, fBN = p.fBN, fTS = p.fTS, fBS = p.fBS;

    velX = fE + fNE + fSE + fTE + fBE;
    velY = fN + fNW + fTN + fBN;
    velZ = fT + fTS + fTW;

    rho = fC + fS + fW + fB + fSW + fBS + fBW + velX + velY + velZ;
    velX = velX - fW - fNW - fSW - fTW - fBW;
    velY = velY + fNE - fS - fSW - fSE - fTS - fBS;
    velZ = velZ + fTN + fTE - fB - fBN - fBS - fBW - fBE;

    const float dir_indep_trm = one_third * rho - 0.5f * (velX * velX + velY * velY + velZ * velZ);

    out[C]  = omega_trm * fC + omega_w0 * (dir_indep_trm);

    out[NW] = omega_trm * fNW + omega_w2 * (dir_indep_trm - (velX - velY) + 1.5f * B200GEO_LBM_SQR(velX - velY));
    out[SE] = omega_trm * fSE + omega_w2 * (dir_indep_trm + (velX - velY) + 1.5f * B200GEO_LBM_SQR(velX - velY));
    out[NE] = omega_trm * fNE + omega_w2 * (dir_indep_trm + (velX + velY) + 1.5f * B200GEO_LBM_SQR(velX + velY));
    out[SW] = omega_trm * fSW + omega_w2 * (dir_indep_trm - (velX + velY) + 1.5f * B200GEO_LBM_SQR(velX + velY));

    out[TW] = omega_trm * fTW + omega_w2 * (dir_indep_trm - (velX - velZ) + 1.5f * B200GEO_LBM_SQR(velX - velZ));
    out[BE] = omega_trm * fBE + omega_w2 * (dir_indep_trm + (velX - velZ) + 1.5f * B200GEO_LBM_SQR(velX - velZ));
    out[TE] = omega_trm * fTE + omega_w2 * (dir_indep_trm + (velX + velZ) + 1.5f * B200GEO_LBM_SQR(velX + velZ));
    out[BW] = omega_trm * fBW + omega_w2 * (dir_indep_trm - (velX + velZ) + 1.5f * B200GEO_LBM_SQR(velX + velZ));

    out[TS] = omega_trm * fTS + omega_w2 * (dir_indep_trm - (velY - velZ) + 1.5f * B200GEO_LBM_SQR(velY - velZ));
    out[BN] = omega_trm * fBN + omega_w2 * (dir_indep_trm + (velY - velZ) + 1.5f * B200GEO_LBM_SQR(velY - velZ));
    out[TN] = omega_trm * fTN + omega_w2 * (dir_indep_trm + (velY + velZ) + 1.5f * B200GEO_LBM_SQR(velY + velZ));
    out[BS] = omega_trm * fBS + omega_w2 * (dir_indep_trm - (velY + velZ) + 1.5f * B200GEO_LBM_SQR(velY + velZ));

    out[N] = omega_trm * fN + omega_w1 * (dir_indep_trm + velY + 1.5f * B200GEO_LBM_SQR(velY));
    out[S] = omega_trm * fS + omega_w1 * (dir_indep_trm - velY + 1.5f * B200GEO_LBM_SQR(velY));
    out[E] = omega_trm * fE + omega_w1 * (dir_indep_trm + velX + 1.5f * B200GEO_LBM_SQR(velX));
    out[W] = omega_trm * fW + omega_w1 * (dir_indep_trm - velX + 1.5f * B200GEO_LBM_SQR(velX));
    out[T] = omega_trm * fT + omega_w1 * (dir_indep_trm + velZ + 1.5f * B200GEO_LBM_SQR(velZ));
    out[B] = omega_trm * fB + omega_w1 * (dir_indep_trm - velZ + 1.5f * B200GEO_LBM_SQR(velZ));
}

#undef B200GEO_LBM_SQR

// asm volatile: neither nvcc nor ptxas may sink these loads below a branch on the cell's state (lbm.cu)
__device__ __forceinline__ float ldg_stream(const float *p)
{
    float v;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

// a cell's neighbourhood in the grid (global memory): member arrays `mstride` elements apart, cell index i
struct GridHood {
    const float *src;
    int64_t i, pitch, plane, mstride;
    template<int X, int Y, int Z, int COMP>
    __device__ __forceinline__ float get() const
    {
        return ldg_stream(src + (int64_t)COMP * mstride + i + X + Y * pitch + Z * plane);
    }
};

}
}

#endif
