/* TEST / MEASUREMENT INFRASTRUCTURE — the D3Q19 BGK cell of models/lbm.h as a plain AoS LibGeoDecomp model with
 * the classic per-cell update(hood, nanoStep) and FixedCoord neighbourhood access, __host__ __device__ like the
 * reference's CUDA-capable cells: the form in which src/examples/latticeboltzmann/main.cpp:62-229 writes it (there
 * in double). Same expression trees as models/lbm.h (generated from it), so it doubles as a cross-check of that
 * model. Used by oracle/ref_cuda_model.cu (the reference's own CUDASimulator on the GPU box) and by
 * tests/facade/generic_test.cu (the generic device path: no binding, no hand kernel). */
#ifndef B200GEO_ORACLE_MODELS_LBM_AOS_H
#define B200GEO_ORACLE_MODELS_LBM_AOS_H

#include <libgeodecomp/misc/apitraits.h>
#include <libgeodecomp/geometry/fixedcoord.h>
#include <libgeodecomp/geometry/stencils.h>

#include <cstring>

#ifndef __CUDACC__
#ifndef __host__
#define __host__
#define __device__
#define B200GEO_LBM_AOS_UNDEF_HD
#endif
#endif

namespace b200models {

using namespace LibGeoDecomp;

class LBMCellAoS
{
public:
    class API : public APITraits::HasFixedCoordsOnlyUpdate,
                public APITraits::HasStencil<Stencils::Moore<3, 1> >,
                public APITraits::HasCubeTopology<3>
    {};

    enum State {LIQUID, WEST_NOSLIP, EAST_NOSLIP, TOP, BOTTOM, NORTH_ACC, SOUTH_NOSLIP};

    __host__ __device__
    inline explicit LBMCellAoS(float v = 1.0f, int s = LIQUID) :
        C(v), N(0), E(0), W(0), S(0), T(0), B(0),
        NW(0), SW(0), NE(0), SE(0),
        TW(0), BW(0), TE(0), BE(0),
        TN(0), BN(0), TS(0), BS(0),
        density(1.0f), velocityX(0), velocityY(0), velocityZ(0),
        state(s)
    {}

#define GET_COMP(X, Y, Z, COMP) hood[FixedCoord<X, Y, Z>()].COMP
#define SQR(X) ((X)*(X))

    template<typename HOOD>
    __host__ __device__
    void update(const HOOD& hood, unsigned /* nanoStep */)
    {
        const int s = GET_COMP(0, 0, 0, state);
        if (s == LIQUID) {
            updateFluid(hood);
            return;
        }
        *this = hood[FixedCoord<0, 0, 0>()];
        switch (s) {
        case WEST_NOSLIP:
            E  = GET_COMP(1, 0,  0, W);
            NE = GET_COMP(1, 1,  0, SW);
            SE = GET_COMP(1,-1,  0, NW);
            TE = GET_COMP(1, 0,  1, BW);
            BE = GET_COMP(1, 0, -1, TW);
            break;
        case EAST_NOSLIP:
            W  = GET_COMP(-1, 0, 0, E);
            NW = GET_COMP(-1, 0, 1, SE);
            SW = GET_COMP(-1,-1, 0, NE);
            TW = GET_COMP(-1, 0, 1, BE);
            BW = GET_COMP(-1, 0,-1, TE);
            break;
        case TOP:
            B  = GET_COMP(0, 0,-1, T);
            BE = GET_COMP(1, 0,-1, TW);
            BW = GET_COMP(-1,0,-1, TE);
            BN = GET_COMP(0, 1,-1, TS);
            BS = GET_COMP(0,-1,-1, TN);
            break;
        case BOTTOM:
            T  = GET_COMP(0, 0, 1, B);
            TE = GET_COMP(1, 0, 1, BW);
            TW = GET_COMP(-1,0, 1, BE);
            TN = GET_COMP(0, 1, 1, BS);
            TS = GET_COMP(0,-1, 1, BN);
            break;
        case NORTH_ACC: {
            const float w_1 = 0.01f;
            S  = GET_COMP(0,-1, 0, N);
            SE = GET_COMP(1,-1, 0, NW) + 6.0f * w_1 * 0.1f;
            SW = GET_COMP(-1,-1,0, NE) - 6.0f * w_1 * 0.1f;
            TS = GET_COMP(0,-1, 1, BN);
            BS = GET_COMP(0,-1,-1, TN);
            break;
        }
        case SOUTH_NOSLIP:
            N  = GET_COMP(0, 1, 0, S);
            NE = GET_COMP(1, 1, 0, SW);
            NW = GET_COMP(-1,1, 0, SE);
            TN = GET_COMP(0, 1, 1, BS);
            BN = GET_COMP(0, 1,-1, TS);
            break;
        }
    }

    template<typename HOOD>
    __host__ __device__
    inline void updateFluid(const HOOD& hood)
    {
        const float omega     = (float)(1.0 / 1.7);
        const float omega_trm = 1.0f - omega;
        const float omega_w0  = (float)(3.0 * 1.0 / 3.0)  * omega;
        const float omega_w1  = (float)(3.0 * 1.0 / 18.0) * omega;
        const float omega_w2  = (float)(3.0 * 1.0 / 36.0) * omega;
        const float one_third = (float)(1.0 / 3.0);
        float velX, velY, velZ;

        velX =
            GET_COMP(-1, 0, 0, E)  + GET_COMP(-1,-1, 0, NE) +
            GET_COMP(-1, 1, 0, SE) + GET_COMP(-1, 0,-1, TE) +
            GET_COMP(-1, 0, 1, BE);
        velY = GET_COMP(0,-1, 0, N) + GET_COMP(1,-1, 0, NW) +
            GET_COMP(0,-1,-1, TN) + GET_COMP(0,-1, 1, BN);
        velZ = GET_COMP(0, 0,-1, T) + GET_COMP(0, 1,-1, TS) +
            GET_COMP(1, 0,-1, TW);

        const float rho =
            GET_COMP(0, 0, 0, C)  + GET_COMP(0, 1, 0, S) +
            GET_COMP(1, 0, 0, W)  + GET_COMP(0, 0, 1, B) +
            GET_COMP(1, 1, 0, SW) + GET_COMP(0, 1, 1, BS) +
            GET_COMP(1, 0, 1, BW) + velX + velY + velZ;
        velX = velX
            - GET_COMP(1, 0, 0, W)  - GET_COMP(1,-1, 0, NW)
            - GET_COMP(1, 1, 0, SW) - GET_COMP(1, 0,-1, TW)
            - GET_COMP(1, 0, 1, BW);
        velY = velY
            + GET_COMP(-1,-1, 0, NE) - GET_COMP(0, 1, 0, S)
            - GET_COMP(1, 1, 0, SW)  - GET_COMP(-1, 1, 0, SE)
            - GET_COMP(0, 1,-1, TS)  - GET_COMP(0, 1, 1, BS);
        velZ = velZ + GET_COMP(0,-1,-1, TN) + GET_COMP(-1, 0,-1, TE) - GET_COMP(0, 0, 1, B)
            - GET_COMP(0,-1, 1, BN) - GET_COMP(0, 1, 1, BS) - GET_COMP(1, 0, 1, BW)
            - GET_COMP(-1, 0, 1, BE);

        density   = rho;
        velocityX = velX;
        velocityY = velY;
        velocityZ = velZ;

        const float dir_indep_trm = one_third * rho - 0.5f * (velX * velX + velY * velY + velZ * velZ);

        C  = omega_trm * GET_COMP(0, 0, 0, C) + omega_w0 * (dir_indep_trm);

        NW = omega_trm * GET_COMP( 1,-1, 0, NW) + omega_w2 * (dir_indep_trm - (velX - velY) + 1.5f * SQR(velX - velY));
        SE = omega_trm * GET_COMP(-1, 1, 0, SE) + omega_w2 * (dir_indep_trm + (velX - velY) + 1.5f * SQR(velX - velY));
        NE = omega_trm * GET_COMP(-1,-1, 0, NE) + omega_w2 * (dir_indep_trm + (velX + velY) + 1.5f * SQR(velX + velY));
        SW = omega_trm * GET_COMP( 1, 1, 0, SW) + omega_w2 * (dir_indep_trm - (velX + velY) + 1.5f * SQR(velX + velY));

        TW = omega_trm * GET_COMP( 1, 0,-1, TW) + omega_w2 * (dir_indep_trm - (velX - velZ) + 1.5f * SQR(velX - velZ));
        BE = omega_trm * GET_COMP(-1, 0, 1, BE) + omega_w2 * (dir_indep_trm + (velX - velZ) + 1.5f * SQR(velX - velZ));
        TE = omega_trm * GET_COMP(-1, 0,-1, TE) + omega_w2 * (dir_indep_trm + (velX + velZ) + 1.5f * SQR(velX + velZ));
        BW = omega_trm * GET_COMP( 1, 0, 1, BW) + omega_w2 * (dir_indep_trm - (velX + velZ) + 1.5f * SQR(velX + velZ));

        TS = omega_trm * GET_COMP(0, 1,-1, TS) + omega_w2 * (dir_indep_trm - (velY - velZ) + 1.5f * SQR(velY - velZ));
        BN = omega_trm * GET_COMP(0,-1, 1, BN) + omega_w2 * (dir_indep_trm + (velY - velZ) + 1.5f * SQR(velY - velZ));
        TN = omega_trm * GET_COMP(0,-1,-1, TN) + omega_w2 * (dir_indep_trm + (velY + velZ) + 1.5f * SQR(velY + velZ));
        BS = omega_trm * GET_COMP(0, 1, 1, BS) + omega_w2 * (dir_indep_trm - (velY + velZ) + 1.5f * SQR(velY + velZ));

        N = omega_trm * GET_COMP(0,-1, 0, N) + omega_w1 * (dir_indep_trm + velY + 1.5f * SQR(velY));
        S = omega_trm * GET_COMP(0, 1, 0, S) + omega_w1 * (dir_indep_trm - velY + 1.5f * SQR(velY));
        E = omega_trm * GET_COMP(-1, 0, 0, E) + omega_w1 * (dir_indep_trm + velX + 1.5f * SQR(velX));
        W = omega_trm * GET_COMP( 1, 0, 0, W) + omega_w1 * (dir_indep_trm - velX + 1.5f * SQR(velX));
        T = omega_trm * GET_COMP(0, 0,-1, T) + omega_w1 * (dir_indep_trm + velZ + 1.5f * SQR(velZ));
        B = omega_trm * GET_COMP(0, 0, 1, B) + omega_w1 * (dir_indep_trm - velZ + 1.5f * SQR(velZ));

        state = LIQUID;
    }

#undef GET_COMP
#undef SQR

    bool operator==(const LBMCellAoS& o) const
    {
        return std::memcmp(this, &o, sizeof(*this)) == 0;
    }

    float C, N, E, W, S, T, B, NW, SW, NE, SE, TW, BW, TE, BE, TN, BN, TS, BS;
    float density, velocityX, velocityY, velocityZ;
    int state;
};

}

#ifdef B200GEO_LBM_AOS_UNDEF_HD
#undef __host__
#undef __device__
#undef B200GEO_LBM_AOS_UNDEF_HD
#endif

#endif
