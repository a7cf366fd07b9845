/* TEST INFRASTRUCTURE — D3Q19 BGK lattice Boltzmann cell in single precision as a
 * SoA + updateLineX LibGeoDecomp model. Arithmetic and boundary states follow
 * src/examples/latticeboltzmann/main.cpp:62-229 (pull scheme, omega = 1/1.7, weights
 * 1/3, 1/18, 1/36, six wall states incl. the accelerated NORTH_ACC lid), member set and
 * SoA registration follow src/testbed/performancetests/main.cpp:1796-1986 (LBMSoACell),
 * with `double` replaced by `float` (BASELINE.json config 4). The example's
 * updateEastNoSlip reads (-1,0,1) for NW (main.cpp:184); kept as is.
 */
#ifndef B200GEO_ORACLE_MODELS_LBM_H
#define B200GEO_ORACLE_MODELS_LBM_H

#include <libgeodecomp/misc/apitraits.h>
#include <libgeodecomp/geometry/fixedcoord.h>
#include <libgeodecomp/geometry/stencils.h>
#include <libflatarray/flat_array.hpp>

namespace b200models {

using namespace LibGeoDecomp;

class LBMCellF
{
public:
    class API : public APITraits::HasFixedCoordsOnlyUpdate,
                public APITraits::HasSoA,
                public APITraits::HasUpdateLineX,
                public APITraits::HasStencil<Stencils::Moore<3, 1> >,
                public APITraits::HasCubeTopology<3>
    {};

    enum State {LIQUID, WEST_NOSLIP, EAST_NOSLIP, TOP, BOTTOM, NORTH_ACC, SOUTH_NOSLIP};

    inline explicit LBMCellF(float v = 1.0f, int s = LIQUID) :
        C(v), N(0), E(0), W(0), S(0), T(0), B(0),
        NW(0), SW(0), NE(0), SE(0),
        TW(0), BW(0), TE(0), BE(0),
        TN(0), BN(0), TS(0), BS(0),
        density(1.0f), velocityX(0), velocityY(0), velocityZ(0),
        state(s)
    {}

#define GET_COMP(X, Y, Z, COMP) hoodOld[FixedCoord<X, Y, Z>()].COMP()
#define SQR(X) ((X)*(X))
#define COPY_COMP(COMP) hoodNew.COMP() = GET_COMP(0, 0, 0, COMP)

    template<typename ACCESSOR1, typename ACCESSOR2>
    static void updateLineX(ACCESSOR1& hoodOld, int indexEnd, ACCESSOR2& hoodNew, int /* nanoStep */)
    {
        for (; hoodOld.index() < indexEnd; ++hoodOld.index(), ++hoodNew.index()) {
            const int s = GET_COMP(0, 0, 0, state);
            if (s == LIQUID) {
                updateFluid(hoodOld, hoodNew);
                continue;
            }

            // *this = neighborhood[FixedCoord<0, 0, 0>()]
            COPY_COMP(C);
            COPY_COMP(N);  COPY_COMP(E);  COPY_COMP(W);  COPY_COMP(S);  COPY_COMP(T);  COPY_COMP(B);
            COPY_COMP(NW); COPY_COMP(SW); COPY_COMP(NE); COPY_COMP(SE);
            COPY_COMP(TW); COPY_COMP(BW); COPY_COMP(TE); COPY_COMP(BE);
            COPY_COMP(TN); COPY_COMP(BN); COPY_COMP(TS); COPY_COMP(BS);
            COPY_COMP(density);
            COPY_COMP(velocityX); COPY_COMP(velocityY); COPY_COMP(velocityZ);

            switch (s) {
            case WEST_NOSLIP:
                hoodNew.E()  = GET_COMP(1, 0,  0, W);
                hoodNew.NE() = GET_COMP(1, 1,  0, SW);
                hoodNew.SE() = GET_COMP(1,-1,  0, NW);
                hoodNew.TE() = GET_COMP(1, 0,  1, BW);
                hoodNew.BE() = GET_COMP(1, 0, -1, TW);
                break;
            case EAST_NOSLIP:
                hoodNew.W()  = GET_COMP(-1, 0, 0, E);
                hoodNew.NW() = GET_COMP(-1, 0, 1, SE);
                hoodNew.SW() = GET_COMP(-1,-1, 0, NE);
                hoodNew.TW() = GET_COMP(-1, 0, 1, BE);
                hoodNew.BW() = GET_COMP(-1, 0,-1, TE);
                break;
            case TOP:
                hoodNew.B()  = GET_COMP(0, 0,-1, T);
                hoodNew.BE() = GET_COMP(1, 0,-1, TW);
                hoodNew.BW() = GET_COMP(-1,0,-1, TE);
                hoodNew.BN() = GET_COMP(0, 1,-1, TS);
                hoodNew.BS() = GET_COMP(0,-1,-1, TN);
                break;
            case BOTTOM:
                hoodNew.T()  = GET_COMP(0, 0, 1, B);
                hoodNew.TE() = GET_COMP(1, 0, 1, BW);
                hoodNew.TW() = GET_COMP(-1,0, 1, BE);
                hoodNew.TN() = GET_COMP(0, 1, 1, BS);
                hoodNew.TS() = GET_COMP(0,-1, 1, BN);
                break;
            case NORTH_ACC: {
                const float w_1 = 0.01f;
                hoodNew.S()  = GET_COMP(0,-1, 0, N);
                hoodNew.SE() = GET_COMP(1,-1, 0, NW) + 6.0f * w_1 * 0.1f;
                hoodNew.SW() = GET_COMP(-1,-1,0, NE) - 6.0f * w_1 * 0.1f;
                hoodNew.TS() = GET_COMP(0,-1, 1, BN);
                hoodNew.BS() = GET_COMP(0,-1,-1, TN);
                break;
            }
            case SOUTH_NOSLIP:
                hoodNew.N()  = GET_COMP(0, 1, 0, S);
                hoodNew.NE() = GET_COMP(1, 1, 0, SW);
                hoodNew.NW() = GET_COMP(-1,1, 0, SE);
                hoodNew.TN() = GET_COMP(0, 1, 1, BS);
                hoodNew.BN() = GET_COMP(0, 1,-1, TS);
                break;
            }
            hoodNew.state() = s;
        }
    }

    template<typename ACCESSOR1, typename ACCESSOR2>
    static inline void updateFluid(ACCESSOR1& hoodOld, ACCESSOR2& hoodNew)
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

        hoodNew.density()   = rho;
        hoodNew.velocityX() = velX;
        hoodNew.velocityY() = velY;
        hoodNew.velocityZ() = velZ;

        const float dir_indep_trm = one_third * rho - 0.5f * (velX * velX + velY * velY + velZ * velZ);

        hoodNew.C()  = omega_trm * GET_COMP(0, 0, 0, C) + omega_w0 * (dir_indep_trm);

        hoodNew.NW() = omega_trm * GET_COMP( 1,-1, 0, NW) + omega_w2 * (dir_indep_trm - (velX - velY) + 1.5f * SQR(velX - velY));
        hoodNew.SE() = omega_trm * GET_COMP(-1, 1, 0, SE) + omega_w2 * (dir_indep_trm + (velX - velY) + 1.5f * SQR(velX - velY));
        hoodNew.NE() = omega_trm * GET_COMP(-1,-1, 0, NE) + omega_w2 * (dir_indep_trm + (velX + velY) + 1.5f * SQR(velX + velY));
        hoodNew.SW() = omega_trm * GET_COMP( 1, 1, 0, SW) + omega_w2 * (dir_indep_trm - (velX + velY) + 1.5f * SQR(velX + velY));

        hoodNew.TW() = omega_trm * GET_COMP( 1, 0,-1, TW) + omega_w2 * (dir_indep_trm - (velX - velZ) + 1.5f * SQR(velX - velZ));
        hoodNew.BE() = omega_trm * GET_COMP(-1, 0, 1, BE) + omega_w2 * (dir_indep_trm + (velX - velZ) + 1.5f * SQR(velX - velZ));
        hoodNew.TE() = omega_trm * GET_COMP(-1, 0,-1, TE) + omega_w2 * (dir_indep_trm + (velX + velZ) + 1.5f * SQR(velX + velZ));
        hoodNew.BW() = omega_trm * GET_COMP( 1, 0, 1, BW) + omega_w2 * (dir_indep_trm - (velX + velZ) + 1.5f * SQR(velX + velZ));

        hoodNew.TS() = omega_trm * GET_COMP(0, 1,-1, TS) + omega_w2 * (dir_indep_trm - (velY - velZ) + 1.5f * SQR(velY - velZ));
        hoodNew.BN() = omega_trm * GET_COMP(0,-1, 1, BN) + omega_w2 * (dir_indep_trm + (velY - velZ) + 1.5f * SQR(velY - velZ));
        hoodNew.TN() = omega_trm * GET_COMP(0,-1,-1, TN) + omega_w2 * (dir_indep_trm + (velY + velZ) + 1.5f * SQR(velY + velZ));
        hoodNew.BS() = omega_trm * GET_COMP(0, 1, 1, BS) + omega_w2 * (dir_indep_trm - (velY + velZ) + 1.5f * SQR(velY + velZ));

        hoodNew.N() = omega_trm * GET_COMP(0,-1, 0, N) + omega_w1 * (dir_indep_trm + velY + 1.5f * SQR(velY));
        hoodNew.S() = omega_trm * GET_COMP(0, 1, 0, S) + omega_w1 * (dir_indep_trm - velY + 1.5f * SQR(velY));
        hoodNew.E() = omega_trm * GET_COMP(-1, 0, 0, E) + omega_w1 * (dir_indep_trm + velX + 1.5f * SQR(velX));
        hoodNew.W() = omega_trm * GET_COMP( 1, 0, 0, W) + omega_w1 * (dir_indep_trm - velX + 1.5f * SQR(velX));
        hoodNew.T() = omega_trm * GET_COMP(0, 0,-1, T) + omega_w1 * (dir_indep_trm + velZ + 1.5f * SQR(velZ));
        hoodNew.B() = omega_trm * GET_COMP(0, 0, 1, B) + omega_w1 * (dir_indep_trm - velZ + 1.5f * SQR(velZ));

        hoodNew.state() = LIQUID;
    }

#undef GET_COMP
#undef SQR
#undef COPY_COMP

    bool operator==(const LBMCellF& o) const
    {
        return C == o.C && N == o.N && E == o.E && W == o.W && S == o.S && T == o.T && B == o.B &&
            NW == o.NW && SW == o.SW && NE == o.NE && SE == o.SE &&
            TW == o.TW && BW == o.BW && TE == o.TE && BE == o.BE &&
            TN == o.TN && BN == o.BN && TS == o.TS && BS == o.BS &&
            density == o.density && velocityX == o.velocityX &&
            velocityY == o.velocityY && velocityZ == o.velocityZ && state == o.state;
    }

    float C, N, E, W, S, T, B, NW, SW, NE, SE, TW, BW, TE, BE, TN, BN, TS, BS;
    float density, velocityX, velocityY, velocityZ;
    int state;
};

}

LIBFLATARRAY_REGISTER_SOA(
    b200models::LBMCellF,
    ((float)(C))((float)(N))((float)(E))((float)(W))((float)(S))((float)(T))((float)(B))
    ((float)(NW))((float)(SW))((float)(NE))((float)(SE))
    ((float)(TW))((float)(BW))((float)(TE))((float)(BE))
    ((float)(TN))((float)(BN))((float)(TS))((float)(BS))
    ((float)(density))((float)(velocityX))((float)(velocityY))((float)(velocityZ))
    ((int)(state)))

#endif
