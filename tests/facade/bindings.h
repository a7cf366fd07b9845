/* The one line per model a user adds to run an existing LibGeoDecomp model on B200Simulator:
 * which hand-written kernel family implements the cell, and its members in registration order. */
#ifndef B200GEO_TESTS_FACADE_BINDINGS_H
#define B200GEO_TESTS_FACADE_BINDINGS_H

#include <libgeodecomp_b200/b200simulator.h>

#include "models/conway.h"
#include "models/jacobi.h"
#include "models/lbm.h"
#include "models/nbody.h"

#define B200_BIND_JACOBI(NAME, KERNEL) \
    B200GEO_BIND_CELL(b200models::NAME, KERNEL, B200GEO_MEMBER_ENTRY(b200models::NAME, temp))

B200_BIND_JACOBI(Jacobi6Cube, B200GEO_KERNEL_JACOBI6)
B200_BIND_JACOBI(Jacobi6Torus, B200GEO_KERNEL_JACOBI6)
B200_BIND_JACOBI(Jacobi7Cube, B200GEO_KERNEL_JACOBI7)
B200_BIND_JACOBI(Jacobi7Torus, B200GEO_KERNEL_JACOBI7)
B200_BIND_JACOBI(Jacobi27Cube, B200GEO_KERNEL_JACOBI27)
B200_BIND_JACOBI(Jacobi27Torus, B200GEO_KERNEL_JACOBI27)

B200GEO_BIND_CELL(b200models::ConwayCube, B200GEO_KERNEL_GOL, B200GEO_MEMBER_ENTRY(b200models::ConwayCube, alive))
B200GEO_BIND_CELL(b200models::ConwayTorus, B200GEO_KERNEL_GOL, B200GEO_MEMBER_ENTRY(b200models::ConwayTorus, alive))

#define B200_LBM_M(M) B200GEO_MEMBER_ENTRY(b200models::LBMCellF, M)
B200GEO_BIND_CELL(b200models::LBMCellF, B200GEO_KERNEL_LBM_D3Q19,
    B200_LBM_M(C) B200_LBM_M(N) B200_LBM_M(E) B200_LBM_M(W) B200_LBM_M(S) B200_LBM_M(T) B200_LBM_M(B)
    B200_LBM_M(NW) B200_LBM_M(SW) B200_LBM_M(NE) B200_LBM_M(SE)
    B200_LBM_M(TW) B200_LBM_M(BW) B200_LBM_M(TE) B200_LBM_M(BE)
    B200_LBM_M(TN) B200_LBM_M(BN) B200_LBM_M(TS) B200_LBM_M(BS)
    B200_LBM_M(density) B200_LBM_M(velocityX) B200_LBM_M(velocityY) B200_LBM_M(velocityZ) B200_LBM_M(state))

/* n-body: where position and velocity live inside the particle, and the model constants */
B200GEO_BIND_PARTICLE(b200models::LJParticle<float>, float, pos, vel,
                      b200models::NBodyParams::dt(), b200models::NBodyParams::cutoff(), b200models::NBodyParams::cellEdge())
B200GEO_BIND_PARTICLE(b200models::LJParticle<double>, double, pos, vel,
                      b200models::NBodyParams::dt(), b200models::NBodyParams::cutoff(), b200models::NBodyParams::cellEdge())

#endif
