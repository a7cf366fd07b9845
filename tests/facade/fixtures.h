/* Shared fixtures of the façade drop-in tests (facade_test.cpp, striping_test.cpp): the user models of
 * oracle/models/*.h with their kernel bindings, seeded Initializers written against the reference's
 * SimpleInitializer, and the CHECK macro. */
#ifndef B200GEO_TESTS_FACADE_FIXTURES_H
#define B200GEO_TESTS_FACADE_FIXTURES_H

#include <libgeodecomp/misc/testcell.h>
#include <libgeodecomp/io/mocksteerer.h>
#include <libgeodecomp/io/mockwriter.h>
#include <libgeodecomp/io/serialbovwriter.h>
#include <libgeodecomp/io/simpleinitializer.h>
#include <libgeodecomp/parallelization/serialsimulator.h>
#include <libgeodecomp/storage/soagrid.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <iterator>

#include "bindings.h"

using namespace LibGeoDecomp;
using namespace b200models;

static int failures = 0;
#define CHECK(COND)                                                                     \
    do {                                                                                \
        if (!(COND)) {                                                                  \
            ++failures;                                                                 \
            std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #COND);              \
        }                                                                               \
    } while (0)

static uint64_t splitmix(uint64_t x)
{
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

static double uniform(uint64_t i)
{
    return (double)(splitmix(i) >> 11) * (1.0 / 9007199254740992.0);
}

template<typename CELL> struct Seed;
template<typename CELL> struct SeedJacobi {
    static CELL make(uint64_t i) { return CELL(uniform(i)); }
    static CELL edge() { return CELL(0.25); }
};
template<> struct Seed<Jacobi6Cube> : SeedJacobi<Jacobi6Cube> {};
template<> struct Seed<Jacobi6Torus> : SeedJacobi<Jacobi6Torus> {};
template<> struct Seed<Jacobi7Cube> : SeedJacobi<Jacobi7Cube> {};
template<> struct Seed<Jacobi7Torus> : SeedJacobi<Jacobi7Torus> {};
template<> struct Seed<Jacobi27Cube> : SeedJacobi<Jacobi27Cube> {};
template<> struct Seed<Jacobi27Torus> : SeedJacobi<Jacobi27Torus> {};
template<> struct Seed<ConwayCube> {
    static ConwayCube make(uint64_t i) { return ConwayCube(uniform(i) < 0.35); }
    static ConwayCube edge() { return ConwayCube(false); }
};
template<> struct Seed<ConwayTorus> {
    static ConwayTorus make(uint64_t i) { return ConwayTorus(uniform(i) < 0.35); }
    static ConwayTorus edge() { return ConwayTorus(false); }
};

template<typename CELL>
class SeededInitializer : public SimpleInitializer<CELL>
{
public:
    typedef typename SimpleInitializer<CELL>::Topology Topology;
    static const int DIM = Topology::DIM;
    using SimpleInitializer<CELL>::gridDimensions;

    SeededInitializer(const Coord<DIM>& dim, unsigned steps) : SimpleInitializer<CELL>(dim, steps) {}

    virtual void grid(GridBase<CELL, DIM> *ret)
    {
        CoordBox<DIM> box = ret->boundingBox();
        ret->setEdge(Seed<CELL>::edge());
        for (typename CoordBox<DIM>::Iterator i = box.begin(); i != box.end(); ++i) {
            ret->set(*i, Seed<CELL>::make(i->toIndex(gridDimensions())));
        }
    }
};

class LBMInitializer : public SimpleInitializer<LBMCellF>
{
public:
    LBMInitializer(const Coord<3>& dim, unsigned steps) : SimpleInitializer<LBMCellF>(dim, steps) {}

    /* walls as src/examples/latticeboltzmann/main.cpp:249-287, rows written with set(Streak, cells) */
    virtual void grid(GridBase<LBMCellF, 3> *ret)
    {
        CoordBox<3> box = ret->boundingBox();
        Coord<3> size = gridDimensions();
        std::vector<LBMCellF> row(box.dimensions.x());
        for (int z = box.origin.z(); z < box.origin.z() + box.dimensions.z(); ++z) {
            for (int y = box.origin.y(); y < box.origin.y() + box.dimensions.y(); ++y) {
                for (int x = 0; x < box.dimensions.x(); ++x) {
                    int gx = box.origin.x() + x;
                    int s = LBMCellF::LIQUID;
                    if (gx == 0) s = LBMCellF::WEST_NOSLIP;
                    if (gx == size.x() - 1) s = LBMCellF::EAST_NOSLIP;
                    if (y == 0) s = LBMCellF::SOUTH_NOSLIP;
                    if (y == size.y() - 1) s = LBMCellF::NORTH_ACC;
                    if (z == 0) s = LBMCellF::BOTTOM;
                    if (z == size.z() - 1) s = LBMCellF::TOP;
                    LBMCellF c(1.0f, s);
                    uint64_t i = Coord<3>(gx, y, z).toIndex(size);
                    c.N = 0.01f * (float)uniform(3 * i);
                    c.TE = 0.01f * (float)uniform(3 * i + 1);
                    c.BS = 0.01f * (float)uniform(3 * i + 2);
                    row[x] = c;
                }
                ret->set(Streak<3>(Coord<3>(box.origin.x(), y, z), box.origin.x() + box.dimensions.x()), row.data());
            }
        }
    }
};


/* n-body: BoxCell<FixedArray<LJParticle<REAL>, 32> > through the reference's SerialSimulator and through
 * B200Simulator (container grid on the device): same occupancy, bit-identical particles */
template<typename REAL>
class ParticleInitializer : public SimpleInitializer<BoxCell<FixedArray<LJParticle<REAL>, NBODY_CAPACITY> > >
{
public:
    typedef BoxCell<FixedArray<LJParticle<REAL>, NBODY_CAPACITY> > Cell;

    ParticleInitializer(const Coord<3>& dim, unsigned steps) : SimpleInitializer<Cell>(dim, steps) {}

    virtual void grid(GridBase<Cell, 3> *ret)
    {
        CoordBox<3> box = ret->boundingBox();
        Coord<3> dim = this->gridDimensions();
        double e = NBodyParams::cellEdge();
        for (CoordBox<3>::Iterator i = box.begin(); i != box.end(); ++i) {
            Cell cell(FloatCoord<3>(i->x() * e, i->y() * e, i->z() * e), FloatCoord<3>(e, e, e));
            uint64_t id = i->toIndex(dim);
            int n = 8 + (int)(splitmix(id) % 9);
            for (int p = 0; p < n; ++p) {
                LJParticle<REAL> particle;
                for (int k = 0; k < 3; ++k) {
                    // well separated sites inside the container, fast enough to cross faces within a few steps
                    double site = ((p >> k) & 1) * 0.5 + (p / 8) * 0.25 + 0.125 + 0.05 * uniform(id * 1000 + p * 8 + k);
                    particle.pos[k] = (REAL)(((*i)[k] + site) * e);
                    particle.vel[k] = (REAL)(16.0 * (uniform(id * 1000 + p * 8 + k + 4) - 0.5));
                }
                cell << particle;
            }
            ret->set(*i, cell);
        }
    }
};

/* The reference's own concrete Writer (io/serialbovwriter.h:20-84 -> BOVOutput::writeGrid, io/bovoutput.h:65-98, which
 * pulls one member row by row through GridBase::saveMemberUnchecked) on the reference's simulator and on SIM: the
 * .bov headers and the .data bricks must be the same bytes. */
static bool sameFile(const std::string& a, const std::string& b)
{
    std::ifstream fa(a.c_str(), std::ios::binary), fb(b.c_str(), std::ios::binary);
    if (!fa.good() || !fb.good()) {
        return false;
    }
    std::vector<char> ca((std::istreambuf_iterator<char>(fa)), std::istreambuf_iterator<char>());
    std::vector<char> cb((std::istreambuf_iterator<char>(fb)), std::istreambuf_iterator<char>());
    return ca.size() > 0 && ca == cb;
}

template<typename SIM>
static void testSerialBOVWriter(SIM& sim, const std::string& tag)
{
    typedef Jacobi7Cube CELL;
    std::string dirRef = "/tmp/b200geo_bov_ref_" + tag, dirDev = "/tmp/b200geo_bov_dev_" + tag;
    {
        SerialSimulator<CELL> ref(new SeededInitializer<CELL>(Coord<3>(20, 7, 12), 6));
        ref.addWriter(new SerialBOVWriter<CELL>(&CELL::temp, dirRef, 3));
        ref.run();
    }
    sim.addWriter(new SerialBOVWriter<CELL>(&CELL::temp, dirDev, 3));
    sim.run();
    int same = 0;
    const char *steps[] = {"00000", "00003", "00006"};
    for (int i = 0; i < 3; ++i) {
        // the header names its own data file, so it differs in that one path; compare the bricks
        same += sameFile(dirRef + "." + steps[i] + ".data", dirDev + "." + steps[i] + ".data");
    }
    CHECK(same == 3);
    std::printf("reference SerialBOVWriter on %s: %d of 3 data bricks byte-identical to the SerialSimulator run\n", tag.c_str(), same);
}

#endif
