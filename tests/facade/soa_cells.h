/* TEST FIXTURES of the generic SoA path (generic_soa_host_test.cpp on the CPU mock, generic_soa_test.cu on the GPU):
 * user-style Struct-of-Arrays cells written against the reference's plugin API only — SoA-signature updateLineX,
 * LIBFLATARRAY_REGISTER_SOA, no B200GEO_BIND_CELL line. The same source runs in the reference's SerialSimulator
 * (SoAGrid + FixedNeighborhoodUpdateFunctor) and on B200Simulator / B200StripingSimulator. Arithmetic is free of
 * a * b + c sites the two compilers could contract differently (both sides are built without contraction anyway).
 * 3-D only: the reference's own SoA dispatch does not compile for 2-D cells (storage/updatefunctor.h:122-123 adds a
 * Coord<2> to the Coord<3> edge radii of SoAGrid). */
#ifndef B200GEO_TESTS_FACADE_SOA_CELLS_H
#define B200GEO_TESTS_FACADE_SOA_CELLS_H

#include <libgeodecomp/geometry/fixedcoord.h>
#include <libgeodecomp/geometry/stencils.h>
#include <libgeodecomp/geometry/topologies.h>
#include <libgeodecomp/misc/apitraits.h>
#include <libflatarray/flat_array.hpp>

#ifndef __host__
#define __host__
#endif
#ifndef __device__
#define __device__
#endif

namespace soacells {

using namespace LibGeoDecomp;

/* 7-point heat cell, one double member: the shape of the testbed's JacobiCellStreakUpdate
 * (src/testbed/performancetests/main.cpp:1276-1395) without its SSE intrinsics */
template<typename TOPOLOGY>
class HeatSoA
{
public:
    class API :
        public APITraits::HasFixedCoordsOnlyUpdate,
        public APITraits::HasUpdateLineX,
        public APITraits::HasStencil<Stencils::VonNeumann<3, 1> >,
        public APITraits::HasTopology<TOPOLOGY>,
        public APITraits::HasSoA
    {};

    __host__ __device__
    inline explicit HeatSoA(double t = 0) : temp(t) {}

    template<typename HOOD_OLD, typename HOOD_NEW>
    __host__ __device__
    static void updateLineX(HOOD_OLD& hoodOld, int indexEnd, HOOD_NEW& hoodNew, int /* nanoStep */)
    {
        for (; hoodOld.index() < indexEnd; ++hoodOld.index(), ++hoodNew.index()) {
            hoodNew.temp() =
                (hoodOld[FixedCoord< 0,  0, -1>()].temp() +
                 hoodOld[FixedCoord< 0, -1,  0>()].temp() +
                 hoodOld[FixedCoord<-1,  0,  0>()].temp() +
                 hoodOld[FixedCoord< 0,  0,  0>()].temp() +
                 hoodOld[FixedCoord< 1,  0,  0>()].temp() +
                 hoodOld[FixedCoord< 0,  1,  0>()].temp() +
                 hoodOld[FixedCoord< 0,  0,  1>()].temp()) * (1.0 / 7.0);
        }
    }

    inline bool operator==(const HeatSoA& o) const { return temp == o.temp; }

    double temp;
};

typedef HeatSoA<Topologies::Cube<3>::Topology> HeatSoACube;
typedef HeatSoA<Topologies::Torus<3>::Topology> HeatSoATorus;

/* members of four widths, an array member, two nano steps, diagonal neighbours (Moore stencil); padding bytes
 * inside the AoS cell (after `flag`) */
template<typename TOPOLOGY>
class MixSoA
{
public:
    class API :
        public APITraits::HasFixedCoordsOnlyUpdate,
        public APITraits::HasUpdateLineX,
        public APITraits::HasStencil<Stencils::Moore<3, 1> >,
        public APITraits::HasTopology<TOPOLOGY>,
        public APITraits::HasNanoSteps<2>,
        public APITraits::HasSoA
    {};

    __host__ __device__
    inline MixSoA() : density(0), count(0), tag(0), flag(0)
    {
        flux[0] = flux[1] = flux[2] = 0;
    }

    template<typename HOOD_OLD, typename HOOD_NEW>
    __host__ __device__
    static void updateLineX(HOOD_OLD& hoodOld, int indexEnd, HOOD_NEW& hoodNew, unsigned nanoStep)
    {
        for (; hoodOld.index() < indexEnd; ++hoodOld.index(), ++hoodNew.index()) {
            double sum =
                hoodOld[FixedCoord<-1, -1, -1>()].density() +
                hoodOld[FixedCoord< 1,  1,  1>()].density() +
                hoodOld[FixedCoord< 1, -1,  0>()].density() +
                hoodOld[FixedCoord< 0,  0,  0>()].density();
            hoodNew.density() = sum * 0.25;
            hoodNew.flux()[0] = hoodOld[FixedCoord<-1, 0, 0>()].flux()[0] * 0.5f + hoodOld[FixedCoord<1, 0, 0>()].flux()[1] * 0.5f;
            hoodNew.flux()[1] = hoodOld[FixedCoord<0, -1, 0>()].flux()[2] - hoodOld[FixedCoord<0, 1, 0>()].flux()[0];
            hoodNew.flux()[2] = hoodOld[FixedCoord<0, 0, -1>()].flux()[1] + (float)nanoStep;
            hoodNew.count() = hoodOld[FixedCoord<0, 0, 0>()].count() + hoodOld[FixedCoord<0, 0, 1>()].flag() + (int)nanoStep;
            hoodNew.tag() = (short)(hoodOld[FixedCoord<1, 0, 0>()].tag() + 1);
            hoodNew.flag() = (char)(hoodOld[FixedCoord<0, 1, -1>()].flag() ^ (hoodOld[FixedCoord<-1, 0, 0>()].count() & 1));
        }
    }

    inline bool operator==(const MixSoA& o) const
    {
        return density == o.density && flux[0] == o.flux[0] && flux[1] == o.flux[1] && flux[2] == o.flux[2] &&
            count == o.count && tag == o.tag && flag == o.flag;
    }

    char flag;
    double density;
    short tag;
    float flux[3];
    int count;
};

typedef MixSoA<Topologies::Cube<3>::Topology> MixSoACube;
typedef MixSoA<Topologies::Torus<3>::Topology> MixSoATorus;

}

LIBFLATARRAY_REGISTER_SOA(soacells::HeatSoACube, ((double)(temp)))
LIBFLATARRAY_REGISTER_SOA(soacells::HeatSoATorus, ((double)(temp)))
LIBFLATARRAY_REGISTER_SOA(soacells::MixSoACube, ((double)(density))((float)(flux)(3))((int)(count))((short)(tag))((char)(flag)))
LIBFLATARRAY_REGISTER_SOA(soacells::MixSoATorus, ((double)(density))((float)(flux)(3))((int)(count))((short)(tag))((char)(flag)))

#endif
