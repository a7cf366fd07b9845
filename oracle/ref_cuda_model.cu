/* TEST / MEASUREMENT INFRASTRUCTURE. The REFERENCE's own GPU path — CUDASimulator<CELL>
 * (src/libgeodecomp/parallelization/cudasimulator.h:298-560, kernel3D :162-236: block 128 x 4 x 1, every thread
 * marching the z extent) — compiled from the headers where they lie under /root/reference with nvcc for sm_100a,
 * running AoS Jacobi cells on the GPU box. This is "the reference's GPU kernel on Blackwell" (SURVEY.md §8d,
 * secondary comparator): what a user gets today by recompiling LibGeoDecomp, and what the hand-written kernels
 * of libb200geo.so are measured against. Nothing in the product links or calls this.
 *
 * usage: lgd_ref_cuda_jacobi [n = 512] [steps = 50] [which = all | 7 | 27 | lbm]
 *        prints one JSON line per cell type (LBM: at most 256^3) */
#include <cuda.h>

#include <libgeodecomp/io/simpleinitializer.h>
#include <libgeodecomp/parallelization/cudasimulator.h>

#include <cstdio>
#include <cstdlib>
#include <string>

#include "models/lbm_aos.h"

using namespace LibGeoDecomp;
using b200models::LBMCellAoS;

class RefJacobi7
{
public:
    class API :
        public APITraits::HasFixedCoordsOnlyUpdate,
        public APITraits::HasStencil<Stencils::VonNeumann<3, 1> >,
        public APITraits::HasCubeTopology<3>
    {};

    __host__ __device__
    explicit RefJacobi7(double temp = 0) : temp(temp)
    {}

    template<typename HOOD>
    __host__ __device__
    void update(const HOOD& hood, int)
    {
        temp = (hood[FixedCoord<0, 0, -1>()].temp + hood[FixedCoord<0, -1, 0>()].temp + hood[FixedCoord<-1, 0, 0>()].temp +
                hood[FixedCoord<0, 0, 0>()].temp + hood[FixedCoord<1, 0, 0>()].temp + hood[FixedCoord<0, 1, 0>()].temp +
                hood[FixedCoord<0, 0, 1>()].temp) * (1.0 / 7.0);
    }

    double temp;
};

class RefJacobi27
{
public:
    class API :
        public APITraits::HasFixedCoordsOnlyUpdate,
        public APITraits::HasStencil<Stencils::Moore<3, 1> >,
        public APITraits::HasCubeTopology<3>
    {};

    __host__ __device__
    explicit RefJacobi27(double temp = 0) : temp(temp)
    {}

#define ROW(Y, Z) ((hood[FixedCoord<-1, Y, Z>()].temp + hood[FixedCoord<0, Y, Z>()].temp) + hood[FixedCoord<1, Y, Z>()].temp)
#define PLANE(Z) ((ROW(-1, Z) + ROW(0, Z)) + ROW(1, Z))
    template<typename HOOD>
    __host__ __device__
    void update(const HOOD& hood, int)
    {
        temp = ((PLANE(-1) + PLANE(0)) + PLANE(1)) * (1.0 / 27.0);
    }
#undef ROW
#undef PLANE

    double temp;
};

template<typename CELL>
class Init : public SimpleInitializer<CELL>
{
public:
    Init(const Coord<3>& dim, unsigned steps) : SimpleInitializer<CELL>(dim, steps) {}

    virtual void grid(GridBase<CELL, 3> *ret)
    {
        CoordBox<3> box = ret->boundingBox();
        std::vector<CELL> row(box.dimensions.x());
        for (CoordBox<3>::StreakIterator i = box.beginStreak(); i != box.endStreak(); ++i) {
            for (std::size_t x = 0; x < row.size(); ++x) {
                row[x] = CELL(0.001 * ((i->origin.x() + x + 3 * i->origin.y() + 7 * i->origin.z()) % 1000));
            }
            ret->set(*i, row.data());
        }
    }
};

/* lid-driven cavity of src/examples/latticeboltzmann/main.cpp:249-287, fluid at rest */
template<>
class Init<LBMCellAoS> : public SimpleInitializer<LBMCellAoS>
{
public:
    Init(const Coord<3>& dim, unsigned steps) : SimpleInitializer<LBMCellAoS>(dim, steps) {}

    virtual void grid(GridBase<LBMCellAoS, 3> *ret)
    {
        CoordBox<3> box = ret->boundingBox();
        Coord<3> size = gridDimensions();
        std::vector<LBMCellAoS> row(box.dimensions.x());
        for (CoordBox<3>::StreakIterator i = box.beginStreak(); i != box.endStreak(); ++i) {
            int y = i->origin.y(), z = i->origin.z();
            for (std::size_t x = 0; x < row.size(); ++x) {
                int gx = i->origin.x() + (int)x;
                int s = LBMCellAoS::LIQUID;
                if (gx == 0) s = LBMCellAoS::WEST_NOSLIP;
                if (gx == size.x() - 1) s = LBMCellAoS::EAST_NOSLIP;
                if (y == 0) s = LBMCellAoS::SOUTH_NOSLIP;
                if (y == size.y() - 1) s = LBMCellAoS::NORTH_ACC;
                if (z == 0) s = LBMCellAoS::BOTTOM;
                if (z == size.z() - 1) s = LBMCellAoS::TOP;
                row[x] = LBMCellAoS(1.0f, s);
            }
            ret->set(*i, row.data());
        }
    }
};

template<typename CELL>
static void bench(const char *name, int n, int steps, int bytesPerUpdate = 16)
{
    Coord<3> dim(n, n, n);
    CUDASimulator<CELL> sim(new Init<CELL>(dim, 3));
    sim.run();                      // initialise, upload, three warm-up steps
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, 0);
    for (int i = 0; i < steps; ++i) {
        sim.step();
    }
    cudaEventRecord(e1, 0);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t e = cudaGetLastError();
    double cells = (double)n * n * n;
    std::printf("{\"impl\": \"reference CUDASimulator (cudasimulator.h, recompiled for sm_100a)\", \"cell\": \"%s\", \"dims\": [%d, %d, %d], "
                "\"steps\": %d, \"ms_per_step\": %.4f, \"glups\": %.2f, \"algorithmic_gbs\": %.0f, \"cuda\": \"%s\"}\n",
                name, n, n, n, steps, ms / steps, 1e-9 * cells * steps / (1e-3 * ms), bytesPerUpdate * 1e-9 * cells * steps / (1e-3 * ms),
                cudaGetErrorString(e));
}

int main(int argc, char **argv)
{
    int n = argc > 1 ? std::atoi(argv[1]) : 512;
    int steps = argc > 2 ? std::atoi(argv[2]) : 50;
    std::string which = argc > 3 ? argv[3] : "all";
    if (which == "all" || which == "7") bench<RefJacobi7>("Jacobi 7-point f64 (AoS, FixedCoord)", n, steps);
    if (which == "all" || which == "27") bench<RefJacobi27>("Jacobi 27-point f64 (AoS, FixedCoord)", n, steps);
    if (which == "all" || which == "lbm")
        bench<LBMCellAoS>("LBM D3Q19 f32 cavity (AoS cell of 96 bytes, FixedCoord)", n > 256 ? 256 : n, steps < 20 ? steps : 20, 152);
    return 0;
}
