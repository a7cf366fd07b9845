/* Drop-in test of the GENERIC device path (include/libgeodecomp_b200/b200generic.h): user cells WITHOUT
 * a B200GEO_BIND_CELL line, their own update() / updateLineX() compiled by nvcc into sm_100a kernels.
 *
 *  1. the reference's own CUDA-simulator suite (parallelization/test/unit/cudasimulatortest.h:13-26,
 *     60-140): TestCell in 1-D/2-D/3-D on Cube and Torus topologies, initialised by the reference's
 *     TestInitializer, checked after EVERY step with the logic of TestWriter / TS_ASSERT_TEST_GRID
 *     (io/testwriter.h:39-62, misc/testhelper.h:99-141) and cell by cell against SerialSimulator;
 *  2. a Game-of-Life cell written like src/examples/gameoflife/main.cpp:25-62 (run-time Coord<2>
 *     neighbour access — not supported by the reference's CUDA path);
 *  3. a two-member heat cell that leaves one member unassigned (sees the new grid's stale value, like
 *     VanillaUpdateFunctor's `gridNew[c].update(...)`);
 *  4. an AoS-signature updateLineX model with NANO_STEPS = 2 on a Torus<3>.
 * Every case runs the reference's SerialSimulator beside B200Simulator in this process and requires
 * bit-identical grids. Compiled HERE (tests/facade/Makefile), run on the GPU box by tests/test_facade_gpu.py. */
#include <cuda.h>

#include <libgeodecomp/io/simpleinitializer.h>
#include <libgeodecomp/io/testinitializer.h>
#include <libgeodecomp/misc/clonable.h>
#include <libgeodecomp/misc/testcell.h>
#include <libgeodecomp/parallelization/serialsimulator.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>

#include <libgeodecomp_b200/b200stripingsimulator.h>
#include <libgeodecomp_b200/b200stepper.h>
#include <libgeodecomp/geometry/partitions/stripingpartition.h>

#include "models/lbm_aos.h"

using namespace LibGeoDecomp;
using b200models::LBMCellAoS;

static int failures = 0;
#define CHECK(COND)                                                                     \
    do {                                                                                \
        if (!(COND)) {                                                                  \
            ++failures;                                                                 \
            std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #COND);              \
        }                                                                               \
    } while (0)

typedef TestCell<1, Stencils::VonNeumann<1, 1>, Topologies::Cube<1>::Topology,
                 TestCellHelpers::EmptyAPI, TestCellHelpers::NoOutput> TestCell1dCube;
typedef TestCell<1, Stencils::Moore<1, 1>, Topologies::Torus<1>::Topology,
                 TestCellHelpers::EmptyAPI, TestCellHelpers::NoOutput> TestCell1dTorus;
typedef TestCell<2, Stencils::VonNeumann<2, 1>, Topologies::Cube<2>::Topology,
                 TestCellHelpers::EmptyAPI, TestCellHelpers::NoOutput> TestCell2dCube;
typedef TestCell<2, Stencils::Moore<2, 1>, Topologies::Torus<2>::Topology,
                 TestCellHelpers::EmptyAPI, TestCellHelpers::NoOutput> TestCell2dTorus;
typedef TestCell<3, Stencils::VonNeumann<3, 1>, Topologies::Cube<3>::Topology,
                 TestCellHelpers::EmptyAPI, TestCellHelpers::NoOutput> TestCell3dCube;
typedef TestCell<3, Stencils::Moore<3, 1>, Topologies::Torus<3>::Topology,
                 TestCellHelpers::EmptyAPI, TestCellHelpers::NoOutput> TestCell3dTorus;
typedef TestCell<3, Stencils::Moore<3, 1>, Topologies::Cube<3>::Topology,
                 TestCellHelpers::EmptyAPI, TestCellHelpers::NoOutput> TestCell3dMooreCube;

/* TestWriter's checks (io/testwriter.h:39-62) without the cxxtest runner: the expected (step, event)
 * sequence, and TS_ASSERT_TEST_GRID on every call */
template<typename CELL>
class CheckingWriter : public Clonable<Writer<CELL>, CheckingWriter<CELL> >
{
public:
    typedef typename Writer<CELL>::GridType GridType;
    static const int DIM = GridType::DIM;
    using Writer<CELL>::NANO_STEPS;

    CheckingWriter(int period, int firstStep, int lastStep, long *badCells, long *badEvents) :
        Clonable<Writer<CELL>, CheckingWriter<CELL> >("", period),
        badCells(badCells),
        badEvents(badEvents)
    {
        expectedSteps.push_back(firstStep);
        expectedEvents.push_back(WRITER_INITIALIZED);
        for (int i = firstStep + period - firstStep % period; i < lastStep; i += period) {
            expectedSteps.push_back(i);
            expectedEvents.push_back(WRITER_STEP_FINISHED);
        }
        expectedSteps.push_back(lastStep);
        expectedEvents.push_back(WRITER_ALL_DONE);
    }

    virtual void stepFinished(const GridType& grid, unsigned step, WriterEvent event)
    {
        if (expectedSteps.empty() || expectedSteps.front() != (int)step || expectedEvents.front() != event) {
            ++*badEvents;
        }
        if (!expectedSteps.empty()) {
            expectedSteps.erase(expectedSteps.begin());
            expectedEvents.erase(expectedEvents.begin());
        }
        unsigned expectedCycle = NANO_STEPS * step;
        if (!grid.getEdge().edgeCell() || !grid.getEdge().valid()) {
            ++*badCells;
        }
        CoordBox<DIM> box = grid.boundingBox();
        std::vector<CELL> row(box.dimensions.x());
        for (typename CoordBox<DIM>::StreakIterator i = box.beginStreak(); i != box.endStreak(); ++i) {
            grid.get(*i, row.data());
            for (std::size_t x = 0; x < row.size(); ++x) {
                if (!row[x].valid() || row[x].edgeCell() || row[x].cycleCounter != expectedCycle) {
                    ++*badCells;
                }
            }
        }
    }

    bool allEventsDone() const
    {
        return expectedSteps.empty() && expectedEvents.empty();
    }

private:
    std::vector<int> expectedSteps;
    std::vector<WriterEvent> expectedEvents;
    long *badCells;
    long *badEvents;
};

template<typename CELL, int DIM>
static long differingCells(const GridBase<CELL, DIM> *a, const GridBase<CELL, DIM> *b)
{
    long bad = 0;
    CoordBox<DIM> box = a->boundingBox();
    if (!(box == b->boundingBox())) {
        return -1;
    }
    std::vector<CELL> ra(box.dimensions.x()), rb(box.dimensions.x());
    for (typename CoordBox<DIM>::StreakIterator i = box.beginStreak(); i != box.endStreak(); ++i) {
        a->get(*i, ra.data());
        b->get(*i, rb.data());
        for (std::size_t x = 0; x < ra.size(); ++x) {
            if (!(ra[x] == rb[x])) {
                ++bad;
            }
        }
    }
    return bad;
}

/* the cases of CUDASimulatorTest::test{1,2,3}d{Cube,Torus} */
template<typename CELL, int DIM>
static void testCellSuite(const char *name, const Coord<DIM>& dim, int steps)
{
    long badCells = 0, badEvents = 0;
    B200Simulator<CELL> sim(new TestInitializer<CELL>(dim, steps));
    CheckingWriter<CELL> *writer = new CheckingWriter<CELL>(1, 0, steps, &badCells, &badEvents);
    sim.addWriter(writer);
    sim.run();
    CHECK(writer->allEventsDone());
    CHECK(badCells == 0);
    CHECK(badEvents == 0);
    CHECK((int)sim.getStep() == steps);

    SerialSimulator<CELL> ref(new TestInitializer<CELL>(dim, steps));
    ref.run();
    long bad = differingCells<CELL, DIM>(ref.getGrid(), sim.getGrid());
    CHECK(bad == 0);
    std::printf("%-22s %d steps x %u nano steps: %ld invalid cells, %ld cells differing from SerialSimulator, events %s\n",
                name, steps, (unsigned)CELL::NANO_STEPS, badCells, bad, badEvents == 0 && writer->allEventsDone() ? "ok" : "WRONG");
}

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

/* Game of Life as a user would write it (src/examples/gameoflife/main.cpp:25-62): default API
 * (Cube<2>, Moore<2,1>), neighbours through run-time Coord<2> */
class LifeCell
{
public:
    class API : public APITraits::HasStencil<Stencils::Moore<2, 1> >
    {};

    __host__ __device__
    explicit LifeCell(bool alive = false) : alive(alive)
    {}

    __host__ __device__
    int countLivingNeighbors(const LifeCell& c) const
    {
        return c.alive ? 1 : 0;
    }

    template<typename COORD_MAP>
    __host__ __device__
    void update(const COORD_MAP& neighborhood, unsigned)
    {
        int livingNeighbors = 0;
        for (int y = -1; y <= 1; ++y) {
            for (int x = -1; x <= 1; ++x) {
                livingNeighbors += countLivingNeighbors(neighborhood[Coord<2>(x, y)]);
            }
        }
        const LifeCell old = neighborhood[Coord<2>(0, 0)];
        livingNeighbors -= old.alive ? 1 : 0;
        alive = old.alive ? (livingNeighbors >= 2 && livingNeighbors <= 3) : (livingNeighbors == 3);
    }

    bool operator==(const LifeCell& o) const
    {
        return alive == o.alive;
    }

    bool alive;
};

/* two members of different type; `hits` is assigned only where the cell is hot, otherwise the new
 * grid's stale value stays (deliberately order dependent: must match the CPU's double buffering) */
class HeatCell
{
public:
    class API :
        public APITraits::HasStencil<Stencils::VonNeumann<2, 1> >,
        public APITraits::HasTorusTopology<2>
    {};

    __host__ __device__
    explicit HeatCell(double temp = 0, long long hits = 0) : temp(temp), hits(hits)
    {}

    template<typename HOOD>
    __host__ __device__
    void update(const HOOD& hood, unsigned nanoStep)
    {
        double sum = hood[FixedCoord< 0, -1>()].temp + hood[FixedCoord<-1, 0>()].temp + hood[FixedCoord<0, 0>()].temp +
            hood[FixedCoord< 1, 0>()].temp + hood[FixedCoord< 0, 1>()].temp;
        temp = sum * 0.2;
        if (temp > 0.55) {
            hits = hood[FixedCoord<0, 0>()].hits + 1 + nanoStep;
        }
    }

    bool operator==(const HeatCell& o) const
    {
        return std::memcmp(&temp, &o.temp, sizeof(temp)) == 0 && hits == o.hits;
    }

    double temp;
    long long hits;
};

/* AoS line signature (src/examples/jacobi3dupdateline/main.cpp:32-42 style), two nano steps that use
 * different formulas, Torus<3>, float members */
class WaveCell
{
public:
    class API :
        public APITraits::HasFixedCoordsOnlyUpdate,
        public APITraits::HasUpdateLineX,
        public APITraits::HasStencil<Stencils::VonNeumann<3, 1> >,
        public APITraits::HasTorusTopology<3>,
        public APITraits::HasNanoSteps<2>
    {};

    __host__ __device__
    explicit WaveCell(float u = 0, float v = 0) : u(u), v(v)
    {}

    template<typename HOOD>
    __host__ __device__
    static void updateLineX(WaveCell *target, long *x, long endX, const HOOD& hood, unsigned nanoStep)
    {
        for (; *x < endX; ++*x) {
            const WaveCell c = hood[FixedCoord<0, 0, 0>()];
            if (nanoStep == 0) {
                float lap = hood[FixedCoord<0, 0, -1>()].u + hood[FixedCoord<0, -1, 0>()].u + hood[FixedCoord<-1, 0, 0>()].u +
                    hood[FixedCoord<1, 0, 0>()].u + hood[FixedCoord<0, 1, 0>()].u + hood[FixedCoord<0, 0, 1>()].u;
                target[*x].v = c.v + 0.1f * (lap - 6.0f * c.u);
                target[*x].u = c.u;
            } else {
                target[*x].u = c.u + 0.5f * c.v;
                target[*x].v = c.v;
            }
        }
    }

    bool operator==(const WaveCell& o) const
    {
        return std::memcmp(this, &o, sizeof(*this)) == 0;
    }

    float u, v;
};

template<typename CELL> struct Seed;
template<> struct Seed<LifeCell> {
    static LifeCell make(uint64_t i) { return LifeCell(uniform(i) < 0.35); }
    static LifeCell edge() { return LifeCell(false); }
};
template<> struct Seed<HeatCell> {
    static HeatCell make(uint64_t i) { return HeatCell(uniform(i), (long long)(i % 7)); }
    static HeatCell edge() { return HeatCell(0.5, -1); }
};
template<> struct Seed<WaveCell> {
    static WaveCell make(uint64_t i) { return WaveCell((float)uniform(2 * i), (float)(uniform(2 * i + 1) - 0.5)); }
    static WaveCell edge() { return WaveCell(0, 0); }
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
        std::vector<CELL> row(box.dimensions.x());
        for (typename CoordBox<DIM>::StreakIterator i = box.beginStreak(); i != box.endStreak(); ++i) {
            Coord<DIM> c = i->origin;
            for (std::size_t x = 0; x < row.size(); ++x, ++c.x()) {
                row[x] = Seed<CELL>::make(c.toIndex(gridDimensions()));
            }
            ret->set(*i, row.data());
        }
    }
};

template<typename CELL, int DIM>
static void compareWithSerialSimulator(const char *name, const Coord<DIM>& dim, unsigned steps)
{
    SerialSimulator<CELL> ref(new SeededInitializer<CELL>(dim, steps));
    B200Simulator<CELL> sim(new SeededInitializer<CELL>(dim, steps));
    ref.run();
    sim.run();
    CHECK(ref.getStep() == steps);
    CHECK(sim.getStep() == steps);
    CHECK(ref.getGrid()->getEdge() == sim.getGrid()->getEdge());
    long bad = differingCells<CELL, DIM>(ref.getGrid(), sim.getGrid());
    CHECK(bad == 0);
    std::printf("%-22s %s: %ld differing cells after %u steps\n", name, bad ? "MISMATCH" : "bit-exact", bad, steps);

    // one step() at a time gives the same grid as run()
    B200Simulator<CELL> single(new SeededInitializer<CELL>(dim, steps));
    for (unsigned i = 0; i < steps; ++i) {
        single.step();
    }
    CHECK(single.getStep() == steps);
    CHECK((differingCells<CELL, DIM>(ref.getGrid(), single.getGrid()) == 0));
}

/* the D3Q19 cavity as a plain 96-byte AoS user cell (oracle/models/lbm_aos.h): 24 words per cell on the device grid */
class CavityInitializer : public SimpleInitializer<LBMCellAoS>
{
public:
    CavityInitializer(const Coord<3>& dim, unsigned steps) : SimpleInitializer<LBMCellAoS>(dim, steps) {}

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
                LBMCellAoS c(1.0f, s);
                uint64_t id = Coord<3>(gx, y, z).toIndex(size);
                c.N = 0.01f * (float)uniform(3 * id);
                c.TE = 0.01f * (float)uniform(3 * id + 1);
                c.BS = 0.01f * (float)uniform(3 * id + 2);
                row[x] = c;
            }
            ret->set(*i, row.data());
        }
    }
};

/* unbound cells on a slab group: the group drives rims / halo copies / interiors and calls the user's update()
 * back for every box (b200geo_group_step_with); slabs round-robin over the GPUs present */
static std::vector<int> devicesFor(int slabs)
{
    int n = b200geo_device_count();
    std::vector<int> ret;
    for (int s = 0; s < slabs; ++s) {
        ret.push_back(n > 0 ? s % n : 0);
    }
    return ret;
}

template<typename CELL, typename INIT, int DIM>
static void compareStriped(const char *name, const Coord<DIM>& dim, unsigned steps, int slabs)
{
    SerialSimulator<CELL> ref(new INIT(dim, steps));
    B200StripingSimulator<CELL> sim(new INIT(dim, steps), devicesFor(slabs));
    ref.run();
    sim.run();
    CHECK(sim.getStep() == steps);
    long bad = differingCells<CELL, DIM>(ref.getGrid(), sim.getGrid());
    CHECK(bad == 0);
    std::pair<unsigned long long, unsigned long long> st = sim.stripedGrid().exchangeStatistics();
    CHECK(slabs == 1 || st.first >= steps);
    std::printf("%-22s on %d slabs: %s (%ld differing cells after %u steps, %llu exchanges)\n", name, slabs,
                bad ? "MISMATCH" : "bit-exact", bad, steps, st.first);
}

/* 7-point Jacobi as a plain user cell (no binding, no hand kernel): what the generic path costs */
class PlainJacobi
{
public:
    class API :
        public APITraits::HasStencil<Stencils::VonNeumann<3, 1> >,
        public APITraits::HasCubeTopology<3>
    {};

    __host__ __device__
    explicit PlainJacobi(double temp = 0) : temp(temp)
    {}

    template<typename HOOD>
    __host__ __device__
    void update(const HOOD& hood, unsigned)
    {
        temp = (hood[FixedCoord<0, 0, -1>()].temp + hood[FixedCoord<0, -1, 0>()].temp + hood[FixedCoord<-1, 0, 0>()].temp +
                hood[FixedCoord<0, 0, 0>()].temp + hood[FixedCoord<1, 0, 0>()].temp + hood[FixedCoord<0, 1, 0>()].temp +
                hood[FixedCoord<0, 0, 1>()].temp) * (1.0 / 7.0);
    }

    bool operator==(const PlainJacobi& o) const
    {
        return std::memcmp(&temp, &o.temp, sizeof(temp)) == 0;
    }

    double temp;
};

template<> struct Seed<PlainJacobi> {
    static PlainJacobi make(uint64_t i) { return PlainJacobi(uniform(i)); }
    static PlainJacobi edge() { return PlainJacobi(0.5); }
};

/* throughput of the generic path (`generic_test --bench`): wall clock around step() calls, grid resident */
template<typename CELL, int DIM, typename INIT = SeededInitializer<CELL> >
static void benchGeneric(const char *name, const Coord<DIM>& dim, unsigned steps, int bytesPerUpdate)
{
    B200Simulator<CELL> sim(new INIT(dim, steps + 3));
    for (int i = 0; i < 3; ++i) {
        sim.step();
    }
    sim.getGrid();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, 0);
    for (unsigned i = 0; i < steps; ++i) {
        sim.step();
    }
    cudaEventRecord(e1, 0);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    double cells = 1;
    for (int d = 0; d < DIM; ++d) {
        cells *= dim[d];
    }
    double glups = 1e-9 * cells * steps * APITraits::SelectNanoSteps<CELL>::VALUE / (1e-3 * ms);
    std::printf("{\"generic\": \"%s\", \"cells\": %.0f, \"steps\": %u, \"ms_per_sweep\": %.4f, \"glups\": %.2f, \"algorithmic_gbs\": %.0f}\n",
                name, cells, steps, ms / (steps * APITraits::SelectNanoSteps<CELL>::VALUE), glups, glups * bytesPerUpdate);
}

/* The reference's stepper test on B200Stepper, with its own TestCell (unbound: the generic device path) —
 * parallelization/nesting/test/parallel_mpi_1/vanillastepperregiontest.h:48-124: a 17 x 12 space, a StripingPartition
 * with ragged weights, this process as rank 1 of 3, ghost zone width 3; after n calls of update1() the cells of
 * innerSet(n) must be valid TestCells of cycle n (TS_ASSERT_TEST_GRID_REGION, misc/testhelper.h:146-180). 3-D likewise. */
template<typename CELL, int DIM>
static void stepperRegionTest(const char *name, const Coord<DIM>& dim, const std::vector<std::size_t>& weights, unsigned ghostZoneWidth)
{
    typedef typename APITraits::SelectTopology<CELL>::Value Topology;
    typedef B200Stepper<CELL> StepperType;
    typedef typename StepperType::GridType GridType;
    typename SharedPtr<TestInitializer<CELL> >::Type init(new TestInitializer<CELL>(dim));
    CoordBox<DIM> rect = init->gridBox();
    typename SharedPtr<Partition<DIM> >::Type partition(new StripingPartition<DIM>(Coord<DIM>(), rect.dimensions, 0, weights));
    typename SharedPtr<AdjacencyManufacturer<DIM> >::Type adjacency(new DummyAdjacencyManufacturer<DIM>);
    typename SharedPtr<PartitionManager<Topology> >::Type manager(new PartitionManager<Topology>());
    manager->resetRegions(adjacency, rect, partition, 1, ghostZoneWidth);
    std::vector<CoordBox<DIM> > boundingBoxes, expandedBoundingBoxes;
    for (int i = 0; i < 3; ++i) {
        Region<DIM> region = partition->getRegion(i);
        boundingBoxes.push_back(region.boundingBox());
        expandedBoundingBoxes.push_back(region.expandWithTopology(ghostZoneWidth, rect.dimensions, Topology()).boundingBox());
    }
    manager->resetGhostZones(boundingBoxes, expandedBoundingBoxes);
    StepperType stepper(manager, init);
    long bad = 0, checked = 0;
    for (unsigned n = 0; n < ghostZoneWidth; ++n) {
        if (n > 0) {
            stepper.update1();
        }
        const GridType& grid = stepper.grid();
        CHECK(grid.getEdge().edgeCell() && grid.getEdge().valid());
        unsigned expectedCycle = n * APITraits::SelectNanoSteps<CELL>::VALUE / APITraits::SelectNanoSteps<CELL>::VALUE;
        const Region<DIM>& region = manager->innerSet(n);
        for (typename Region<DIM>::Iterator i = region.begin(); i != region.end(); ++i) {
            CELL cell = grid.get(*i);
            bad += !(cell.valid() && !cell.edgeCell() && cell.cycleCounter == n);
            ++checked;
        }
        (void)expectedCycle;
    }
    CHECK(bad == 0 && checked > 0);
    std::printf("B200Stepper<%s> as rank 1 of 3, ghost zone width %u: %ld of %ld inner set cells wrong after update1() x 0..%u, %zu launches\n",
                name, ghostZoneWidth, bad, checked, ghostZoneWidth - 1, stepper.launchCount());
}

int main(int argc, char **argv)
{
    if (argc > 1 && std::string(argv[1]) == "--bench") {
        try {
            benchGeneric<PlainJacobi, 3>("PlainJacobi 7-point f64 384^3 (user update(), FixedCoord)", Coord<3>(384, 384, 384), 30, 16);
            benchGeneric<LifeCell, 2>("LifeCell 8192^2 (user update(), run-time Coord<2>)", Coord<2>(8192, 8192), 50, 2);
            benchGeneric<WaveCell, 3>("WaveCell 2 x f32 256^3 Torus (user updateLineX, 2 nano steps)", Coord<3>(256, 256, 256), 30, 16);
            benchGeneric<LBMCellAoS, 3, CavityInitializer>("LBM D3Q19 f32 cavity 256^3 (96-byte AoS user cell, update())", Coord<3>(256, 256, 256), 20, 152);
        } catch (const std::exception& e) {
            std::printf("FAILED with exception: %s\n", e.what());
            return 2;
        }
        return 0;
    }
    try {
        CHECK(B200KernelBinding<LifeCell>::kernel() == B200GEO_KERNEL_GENERIC);
        CHECK(B200Generic::Words<LifeCell>::N == 1 && B200Generic::Words<LifeCell>::W == 1);
        CHECK(B200Generic::Words<HeatCell>::N == 2 && B200Generic::Words<HeatCell>::W == 8);
        CHECK(B200Generic::Words<TestCell3dCube>::W == 4);

        testCellSuite<TestCell1dCube, 1>("TestCell 1d Cube", Coord<1>(777), 33);
        testCellSuite<TestCell1dTorus, 1>("TestCell 1d Torus", Coord<1>(666), 33);
        testCellSuite<TestCell2dCube, 2>("TestCell 2d Cube", Coord<2>(64, 32), 33);
        testCellSuite<TestCell2dTorus, 2>("TestCell 2d Torus", Coord<2>(123, 77), 15);
        testCellSuite<TestCell3dCube, 3>("TestCell 3d Cube", Coord<3>(50, 20, 10), 5);
        testCellSuite<TestCell3dTorus, 3>("TestCell 3d Torus", Coord<3>(30, 20, 10), 5);
        testCellSuite<TestCell3dMooreCube, 3>("TestCell 3d Moore Cube", Coord<3>(13, 12, 11), 4);

        {
            std::vector<std::size_t> w2(3), w3(3);
            // the reference test's weights give rank 1 two rows: with ghost zone width 3 all its inner sets are empty
            // and no cell is checked; here rank 1 is 10 rows / planes thick
            w2[0] = 4 * 17 + 7; w2[1] = 10 * 17 - 1; w2[2] = 20 * 17 - w2[0] - w2[1];
            stepperRegionTest<TestCell2dCube, 2>("TestCell 2d Cube", Coord<2>(17, 20), w2, 3);
            w3[0] = 3 * 66 + 2 * 11 + 5; w3[1] = 10 * 66 + 3 * 11 - 2; w3[2] = 11 * 6 * 18 - w3[0] - w3[1];
            stepperRegionTest<TestCell3dCube, 3>("TestCell 3d Cube", Coord<3>(11, 6, 18), w3, 2);
            stepperRegionTest<TestCell3dMooreCube, 3>("TestCell 3d Moore Cube", Coord<3>(11, 6, 18), w3, 3);
        }

        compareWithSerialSimulator<LifeCell, 2>("LifeCell (Coord<2>)", Coord<2>(150, 67), 30);
        compareWithSerialSimulator<HeatCell, 2>("HeatCell (stale member)", Coord<2>(97, 41), 23);
        compareWithSerialSimulator<WaveCell, 3>("WaveCell (updateLineX)", Coord<3>(40, 9, 7), 12);
        compareWithSerialSimulator<PlainJacobi, 3>("PlainJacobi (7-point)", Coord<3>(33, 10, 9), 9);

        compareStriped<TestCell2dCube, TestInitializer<TestCell2dCube>, 2>("TestCell 2d Cube", Coord<2>(64, 32), 9, 3);
        compareStriped<TestCell2dTorus, TestInitializer<TestCell2dTorus>, 2>("TestCell 2d Torus", Coord<2>(50, 21), 7, 2);
        compareStriped<TestCell3dCube, TestInitializer<TestCell3dCube>, 3>("TestCell 3d Cube", Coord<3>(20, 10, 12), 4, 4);
        compareStriped<TestCell3dTorus, TestInitializer<TestCell3dTorus>, 3>("TestCell 3d Torus", Coord<3>(13, 12, 11), 4, 3);
        compareStriped<TestCell3dMooreCube, TestInitializer<TestCell3dMooreCube>, 3>("TestCell 3d Moore Cube", Coord<3>(13, 12, 3), 3, 3);
        compareStriped<LifeCell, SeededInitializer<LifeCell>, 2>("LifeCell (Coord<2>)", Coord<2>(150, 67), 30, 4);
        compareStriped<HeatCell, SeededInitializer<HeatCell>, 2>("HeatCell (stale member)", Coord<2>(97, 41), 23, 2);
        compareStriped<WaveCell, SeededInitializer<WaveCell>, 3>("WaveCell (updateLineX)", Coord<3>(40, 9, 14), 12, 3);
        compareStriped<PlainJacobi, SeededInitializer<PlainJacobi>, 3>("PlainJacobi (7-point)", Coord<3>(33, 10, 9), 9, 1);
        compareStriped<LBMCellAoS, CavityInitializer, 3>("LBMCellAoS (96-byte cell)", Coord<3>(18, 12, 10), 15, 1);
        compareStriped<LBMCellAoS, CavityInitializer, 3>("LBMCellAoS (96-byte cell)", Coord<3>(18, 12, 10), 15, 3);
    } catch (const std::exception& e) {
        std::printf("FAILED with exception: %s\n", e.what());
        return 2;
    }
    if (failures) {
        std::printf("%d check(s) FAILED\n", failures);
        return 1;
    }
    std::printf("generic_test: all checks passed\n");
    return 0;
}
