/* TEST INFRASTRUCTURE — user-side model definitions written against the unchanged
 * LibGeoDecomp plugin API (Cell + Cell::API traits, misc/apitraits.h:191-1168).
 * They are compiled (a) against the reference's own SerialSimulator/OpenMPSimulator to
 * form the oracle (oracle/_ref/), and (b) against B200Simulator (include/libgeodecomp_b200/)
 * to show that the same model source drops in unchanged.
 *
 * Jacobi6*   : 6-point mean, exactly the arithmetic of src/examples/jacobi3d/main.cpp:30-39
 *              (AoS, HasFixedCoordsOnlyUpdate, per-cell update()).
 * Jacobi7*   : 7-point mean as a SoA + updateLineX cell, add order z-,y-,x-,centre,x+,y+,z+
 *              (the order of src/testbed/performancetests/main.cpp:1292-1299 with its
 *              duplicated <1,0,0> read replaced by the intended <-1,0,0>).
 * Jacobi27*  : 27-point mean over Moore<3,1>; rows are summed first ((W+C)+E), then the three
 *              rows of a plane, then the three planes, then one multiply by 1/27.
 * Each model exists for Cube (constant edge cell) and Torus (periodic) topologies.
 */
#ifndef B200GEO_ORACLE_MODELS_JACOBI_H
#define B200GEO_ORACLE_MODELS_JACOBI_H

#include <libgeodecomp/misc/apitraits.h>
#include <libgeodecomp/geometry/fixedcoord.h>
#include <libgeodecomp/geometry/stencils.h>
#include <libflatarray/flat_array.hpp>

namespace b200models {

using namespace LibGeoDecomp;

#define B200_DEFINE_JACOBI6(NAME, TOPOLOGY_TRAIT)                               \
class NAME                                                                      \
{                                                                               \
public:                                                                         \
    class API :                                                                 \
        public APITraits::HasFixedCoordsOnlyUpdate,                             \
        public APITraits::HasStencil<Stencils::VonNeumann<3, 1> >,              \
        public TOPOLOGY_TRAIT                                                   \
    {};                                                                         \
                                                                                \
    inline explicit NAME(double v = 0) : temp(v) {}                             \
                                                                                \
    template<typename HOOD>                                                     \
    void update(const HOOD& hood, unsigned /* nanoStep */)                      \
    {                                                                           \
        temp = (hood[FixedCoord< 0,  0, -1>()].temp +                           \
                hood[FixedCoord< 0, -1,  0>()].temp +                           \
                hood[FixedCoord<-1,  0,  0>()].temp +                           \
                hood[FixedCoord< 1,  0,  0>()].temp +                           \
                hood[FixedCoord< 0,  1,  0>()].temp +                           \
                hood[FixedCoord< 0,  0,  1>()].temp) * (1.0 / 6.0);             \
    }                                                                           \
                                                                                \
    bool operator==(const NAME& o) const { return temp == o.temp; }             \
                                                                                \
    double temp;                                                                \
};

#define B200_DEFINE_JACOBI7(NAME, TOPOLOGY_TRAIT)                               \
class NAME                                                                      \
{                                                                               \
public:                                                                         \
    class API :                                                                 \
        public APITraits::HasFixedCoordsOnlyUpdate,                             \
        public APITraits::HasUpdateLineX,                                       \
        public APITraits::HasStencil<Stencils::VonNeumann<3, 1> >,              \
        public TOPOLOGY_TRAIT,                                                  \
        public APITraits::HasSoA                                                \
    {};                                                                         \
                                                                                \
    inline explicit NAME(double v = 0) : temp(v) {}                             \
                                                                                \
    template<typename HOOD_OLD, typename HOOD_NEW>                              \
    static void updateLineX(HOOD_OLD& hoodOld, int indexEnd,                    \
                            HOOD_NEW& hoodNew, int /* nanoStep */)              \
    {                                                                           \
        for (; hoodOld.index() < indexEnd; ++hoodOld.index(), ++hoodNew.index()) { \
            hoodNew.temp() =                                                    \
                (hoodOld[FixedCoord< 0,  0, -1>()].temp() +                     \
                 hoodOld[FixedCoord< 0, -1,  0>()].temp() +                     \
                 hoodOld[FixedCoord<-1,  0,  0>()].temp() +                     \
                 hoodOld[FixedCoord< 0,  0,  0>()].temp() +                     \
                 hoodOld[FixedCoord< 1,  0,  0>()].temp() +                     \
                 hoodOld[FixedCoord< 0,  1,  0>()].temp() +                     \
                 hoodOld[FixedCoord< 0,  0,  1>()].temp()) * (1.0 / 7.0);       \
        }                                                                       \
    }                                                                           \
                                                                                \
    bool operator==(const NAME& o) const { return temp == o.temp; }             \
                                                                                \
    double temp;                                                                \
};

#define B200_J27_ROW(Y, Z)                                                      \
    ((hoodOld[FixedCoord<-1, Y, Z>()].temp() +                                  \
      hoodOld[FixedCoord< 0, Y, Z>()].temp()) +                                 \
      hoodOld[FixedCoord< 1, Y, Z>()].temp())
#define B200_J27_PLANE(Z)                                                       \
    ((B200_J27_ROW(-1, Z) + B200_J27_ROW(0, Z)) + B200_J27_ROW(1, Z))

#define B200_DEFINE_JACOBI27(NAME, TOPOLOGY_TRAIT)                              \
class NAME                                                                      \
{                                                                               \
public:                                                                         \
    class API :                                                                 \
        public APITraits::HasFixedCoordsOnlyUpdate,                             \
        public APITraits::HasUpdateLineX,                                       \
        public APITraits::HasStencil<Stencils::Moore<3, 1> >,                   \
        public TOPOLOGY_TRAIT,                                                  \
        public APITraits::HasSoA                                                \
    {};                                                                         \
                                                                                \
    inline explicit NAME(double v = 0) : temp(v) {}                             \
                                                                                \
    template<typename HOOD_OLD, typename HOOD_NEW>                              \
    static void updateLineX(HOOD_OLD& hoodOld, int indexEnd,                    \
                            HOOD_NEW& hoodNew, int /* nanoStep */)              \
    {                                                                           \
        for (; hoodOld.index() < indexEnd; ++hoodOld.index(), ++hoodNew.index()) { \
            hoodNew.temp() =                                                    \
                ((B200_J27_PLANE(-1) + B200_J27_PLANE(0)) + B200_J27_PLANE(1))  \
                * (1.0 / 27.0);                                                 \
        }                                                                       \
    }                                                                           \
                                                                                \
    bool operator==(const NAME& o) const { return temp == o.temp; }             \
                                                                                \
    double temp;                                                                \
};

B200_DEFINE_JACOBI6(Jacobi6Cube,   APITraits::HasCubeTopology<3>)
B200_DEFINE_JACOBI6(Jacobi6Torus,  APITraits::HasTorusTopology<3>)
B200_DEFINE_JACOBI7(Jacobi7Cube,   APITraits::HasCubeTopology<3>)
B200_DEFINE_JACOBI7(Jacobi7Torus,  APITraits::HasTorusTopology<3>)
B200_DEFINE_JACOBI27(Jacobi27Cube,  APITraits::HasCubeTopology<3>)
B200_DEFINE_JACOBI27(Jacobi27Torus, APITraits::HasTorusTopology<3>)

}

LIBFLATARRAY_REGISTER_SOA(b200models::Jacobi7Cube,   ((double)(temp)))
LIBFLATARRAY_REGISTER_SOA(b200models::Jacobi7Torus,  ((double)(temp)))
LIBFLATARRAY_REGISTER_SOA(b200models::Jacobi27Cube,  ((double)(temp)))
LIBFLATARRAY_REGISTER_SOA(b200models::Jacobi27Torus, ((double)(temp)))

#endif
