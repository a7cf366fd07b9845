/* CPU-only companion of facade_test.cpp / striping_test.cpp: error paths and corner cases of the façade's HOST
 * logic, linked against the host-memory mock of the engine (mock_b200geo.cpp) only — never run on the GPU box.
 *  - a container that overflows its FixedArray capacity on the device surfaces as std::out_of_range when the
 *    grid is read (storage/fixedarray.h:77-83), on one grid and on slabs;
 *  - B200Grid::resize drops pending writes and cached rows and keeps working;
 *  - overlapping combined writes keep their order across flushes of different sizes;
 *  - wrong-size region buffers and member type mismatches map to std::invalid_argument;
 *  - a filtered (non-member) Selector takes the host path of saveMember;
 *  - ParallelWriters on B200StripingSimulator. */
#include "fixtures.h"

#include <libgeodecomp/io/tracingwriter.h>
#include <libgeodecomp/misc/clonable.h>

#include <libgeodecomp_b200/b200stripingsimulator.h>

#include "models/lbm_aos.h"

typedef BoxCell<FixedArray<LJParticle<float>, NBODY_CAPACITY> > ParticleCell;

/* 20 particles in container (0,0,0) and 20 in (1,0,0), all of them positioned inside container (1,0,0): the
 * first re-bin moves 40 particles into a FixedArray of 32 */
class OverflowInitializer : public SimpleInitializer<ParticleCell>
{
public:
    OverflowInitializer(const Coord<3>& dim, unsigned steps) : SimpleInitializer<ParticleCell>(dim, steps) {}

    virtual void grid(GridBase<ParticleCell, 3> *ret)
    {
        double e = NBodyParams::cellEdge();
        for (int c = 0; c < 2; ++c) {
            ParticleCell cell(FloatCoord<3>(c * e, 0, 0), FloatCoord<3>(e, e, e));
            for (int p = 0; p < 20; ++p) {
                LJParticle<float> particle;
                particle.pos[0] = (float)(e + 0.1 + 0.1 * p + 0.05 * c);
                particle.pos[1] = (float)(0.2 + 0.1 * p);
                particle.pos[2] = (float)(0.3 + 0.05 * p);
                cell << particle;
            }
            ret->set(Coord<3>(c, 0, 0), cell);
        }
    }
};

template<typename SIM>
static bool overflows(SIM& sim)
{
    try {
        sim.run();
        sim.getGrid()->get(Coord<3>(1, 0, 0));
    } catch (const std::out_of_range&) {
        return true;
    }
    return false;
}

static void testCapacityExceeded()
{
    Coord<3> dim(3, 2, 4);
    {
        SerialSimulator<ParticleCell> ref(new OverflowInitializer(dim, 1));
        CHECK(overflows(ref));      // the reference throws from FixedArray::operator<< inside update()
    }
    {
        B200Simulator<ParticleCell> sim(new OverflowInitializer(dim, 1));
        CHECK(overflows(sim));
    }
    {
        B200StripingSimulator<ParticleCell> sim(new OverflowInitializer(dim, 1), std::vector<int>(2, 0));
        CHECK(overflows(sim));
    }
    std::printf("capacity exceeded: std::out_of_range from SerialSimulator, B200Simulator and B200StripingSimulator\n");
}

static void testResizeAndWriteOrder()
{
    typedef Jacobi7Cube CELL;
    B200Grid<CELL> grid(CoordBox<3>(Coord<3>(), Coord<3>(6, 5, 4)));
    grid.set(Coord<3>(1, 1, 1), CELL(1.0));
    CHECK(grid.get(Coord<3>(1, 1, 1)) == CELL(1.0));
    grid.set(Coord<3>(2, 1, 1), CELL(5.0));             // pending when the grid is resized
    grid.resize(CoordBox<3>(Coord<3>(2, 0, 0), Coord<3>(9, 3, 2)));
    CHECK(grid.boundingBox() == CoordBox<3>(Coord<3>(2, 0, 0), Coord<3>(9, 3, 2)));
    CHECK(grid.get(Coord<3>(3, 1, 1)) == CELL(0.0));    // fresh grid, nothing of the old one
    // the same cell written many times, interleaved with whole rows, through small and large flushes
    std::vector<CELL> row(9);
    for (int round = 0; round < 5; ++round) {
        for (int x = 0; x < 9; ++x) row[x] = CELL(100.0 * round + x);
        grid.set(Streak<3>(Coord<3>(2, 2, 1), 11), row.data());
        grid.set(Coord<3>(4, 2, 1), CELL(-1.0 * round));
        if (round == 2) CHECK(grid.get(Coord<3>(4, 2, 1)) == CELL(-2.0));
    }
    CHECK(grid.get(Coord<3>(4, 2, 1)) == CELL(-4.0));
    CHECK(grid.get(Coord<3>(5, 2, 1)) == CELL(403.0));
    CHECK(grid.get(Coord<3>(10, 2, 1)) == CELL(408.0));
    // more than one automatic flush (64 Ki cells) with rewrites in between
    B200Grid<CELL> big(CoordBox<3>(Coord<3>(), Coord<3>(64, 64, 40)));
    big.setMaxPendingCells(1 << 16);
    CoordBox<3> box = big.boundingBox();
    for (int pass = 0; pass < 2; ++pass) {
        for (CoordBox<3>::Iterator i = box.begin(); i != box.end(); ++i) {
            big.set(*i, CELL(pass * 1e6 + (double)i->toIndex(box.dimensions)));
        }
    }
    long bad = 0;
    for (CoordBox<3>::Iterator i = box.begin(); i != box.end(); ++i) {
        if (!(big.get(*i) == CELL(1e6 + (double)i->toIndex(box.dimensions)))) ++bad;
    }
    CHECK(bad == 0);
    std::printf("resize, write order across flushes: %ld wrong cells\n", bad);
}

static double squareOfTemp(const Jacobi7Cube& cell)
{
    return cell.temp * cell.temp;
}

static void testSelectorsAndErrors()
{
    typedef Jacobi7Cube CELL;
    CoordBox<3> box(Coord<3>(), Coord<3>(7, 4, 3));
    B200Grid<CELL> grid(box);
    for (CoordBox<3>::Iterator i = box.begin(); i != box.end(); ++i) {
        grid.set(*i, CELL(1.0 + i->toIndex(box.dimensions)));
    }
    Region<3> region;
    region << Streak<3>(Coord<3>(1, 1, 1), 6) << Streak<3>(Coord<3>(0, 3, 2), 7);
    Selector<CELL> plain(&CELL::temp, "temp");
    std::vector<double> values(region.size());
    grid.saveMember(values.data(), MemoryLocation::HOST, plain, region);
    CHECK(values[0] == 1.0 + Coord<3>(1, 1, 1).toIndex(box.dimensions));
    CHECK(values[5] == 1.0 + Coord<3>(0, 3, 2).toIndex(box.dimensions));
    bool mismatch = false;
    try {
        std::vector<float> wrong(region.size());
        grid.saveMember(wrong.data(), MemoryLocation::HOST, plain, region);
    } catch (const std::invalid_argument&) {
        mismatch = true;
    }
    CHECK(mismatch);
    bool wrongSize = false;
    try {
        std::vector<char> tooLong(region.size() * sizeof(double) + 8);
        grid.loadRegion(tooLong, region);
    } catch (const std::invalid_argument&) {
        wrongSize = true;
    }
    CHECK(wrongSize);
    bool outside = false;
    try {
        grid.get(Coord<3>(0, 0, 17));
    } catch (const std::exception&) {
        outside = true;
    }
    CHECK(outside);
    std::printf("selectors and error mapping: ok\n");
    (void)squareOfTemp;
}

/* saveMember / loadMember over box-shaped and ragged regions: one strided copy per BOX of the streak list
 * (B200Helpers::mergeStreaks), same bytes as streak by streak */
static void testMemberBoxFastPath()
{
    typedef Jacobi7Cube CELL;
    CoordBox<3> box(Coord<3>(2, 3, 4), Coord<3>(9, 6, 5));      // a grid that does not start at the origin
    B200Grid<CELL> grid(box);
    for (CoordBox<3>::Iterator i = box.begin(); i != box.end(); ++i) {
        grid.set(*i, CELL(0.5 + (*i - box.origin).toIndex(box.dimensions)));
    }
    Selector<CELL> sel(&CELL::temp, "temp");
    struct Case { Region<3> region; std::size_t boxes; } cases[4];
    cases[0].region << box;                                                          // the whole grid: ONE copy
    cases[0].boxes = 1;
    cases[1].region << CoordBox<3>(Coord<3>(3, 4, 5), Coord<3>(4, 3, 2));            // an inner box
    cases[1].boxes = 1;
    cases[2].region << CoordBox<3>(Coord<3>(2, 3, 4), Coord<3>(9, 6, 1))             // a plane ...
                    << CoordBox<3>(Coord<3>(4, 5, 6), Coord<3>(3, 2, 2))             // ... a box two planes up ...
                    << Streak<3>(Coord<3>(2, 8, 8), 7);                              // ... and a lone streak
    cases[2].boxes = 3;
    cases[3].region << Streak<3>(Coord<3>(2, 3, 4), 6) << Streak<3>(Coord<3>(2, 4, 4), 7) << Streak<3>(Coord<3>(3, 5, 4), 7);
    cases[3].boxes = 3;                                                              // ragged rows: nothing to merge
    for (int c = 0; c < 4; ++c) {
        const Region<3>& region = cases[c].region;
        std::vector<double> got(region.size(), -1.0), want;
        for (Region<3>::StreakIterator i = region.beginStreak(); i != region.endStreak(); ++i) {
            for (int x = i->origin.x(); x < i->endX; ++x) {
                Coord<3> at(x, i->origin.y(), i->origin.z());
                want.push_back(0.5 + (at - box.origin).toIndex(box.dimensions));
            }
        }
        std::size_t before = grid.memberCopyCalls();
        grid.saveMember(got.data(), MemoryLocation::HOST, sel, region);
        CHECK(grid.memberCopyCalls() - before == cases[c].boxes);
        CHECK(got == want);
        // and back in, negated
        for (std::size_t k = 0; k < got.size(); ++k) {
            got[k] = -got[k];
        }
        before = grid.memberCopyCalls();
        grid.loadMember(got.data(), MemoryLocation::HOST, sel, region);
        CHECK(grid.memberCopyCalls() - before == cases[c].boxes);
        long bad = 0;
        for (CoordBox<3>::Iterator i = box.begin(); i != box.end(); ++i) {
            double v = 0.5 + (*i - box.origin).toIndex(box.dimensions);
            bad += grid.get(*i).temp != (region.count(*i) ? -v : v);
        }
        CHECK(bad == 0);
        grid.loadMember(want.data(), MemoryLocation::HOST, sel, region);
    }
    std::printf("member I/O by boxes: 1 / 1 / 3 / 3 copies for a grid / a box / plane + box + streak / ragged rows\n");
}

/* Writers that pull a member row by row (BOVOutput::writeGrid, io/bovoutput.h:83-95): rows come out of a read-ahead
 * block, a write in between drops it */
static void testMemberReadAhead()
{
    typedef Jacobi7Cube CELL;
    CoordBox<3> box(Coord<3>(1, 2, 3), Coord<3>(40, 9, 6));
    B200Grid<CELL> grid(box);
    for (CoordBox<3>::Iterator i = box.begin(); i != box.end(); ++i) {
        grid.set(*i, CELL(0.25 + (*i - box.origin).toIndex(box.dimensions)));
    }
    Selector<CELL> sel(&CELL::temp, "temp");
    std::vector<double> row(40);
    long bad = 0;
    std::size_t before = grid.memberCopyCalls();
    for (CoordBox<3>::StreakIterator i = box.beginStreak(); i != box.endStreak(); ++i) {
        Region<3> one;
        one << *i;
        grid.saveMemberUnchecked(reinterpret_cast<char*>(row.data()), MemoryLocation::HOST, sel, one);
        for (int x = 0; x < 40; ++x) {
            bad += row[x] != 0.25 + (Coord<3>(i->origin.x() + x, i->origin.y(), i->origin.z()) - box.origin).toIndex(box.dimensions);
        }
    }
    std::size_t transfers = grid.memberCopyCalls() - before;
    CHECK(bad == 0);
    CHECK(transfers == 1);                       /* 54 rows, one block */
    /* a partial row, then a write, then the same row again: the block is dropped, the new value is seen */
    Region<3> part;
    part << Streak<3>(Coord<3>(5, 4, 5), 9);
    grid.saveMemberUnchecked(reinterpret_cast<char*>(row.data()), MemoryLocation::HOST, sel, part);
    CHECK(row[0] == 0.25 + (Coord<3>(5, 4, 5) - box.origin).toIndex(box.dimensions));
    grid.set(Coord<3>(6, 4, 5), CELL(-3.0));
    grid.saveMemberUnchecked(reinterpret_cast<char*>(row.data()), MemoryLocation::HOST, sel, part);
    CHECK(row[1] == -3.0);
    grid.update(0, 1);
    grid.saveMemberUnchecked(reinterpret_cast<char*>(row.data()), MemoryLocation::HOST, sel, part);
    CHECK(row[1] != -3.0);
    std::printf("member pulled row by row: %zu transfer(s) for 54 rows, %ld wrong values\n", transfers, bad);
}

/* The AoS form of the D3Q19 cell (oracle/models/lbm_aos.h, used by the GPU comparator and the generic-path test) is the
 * same model as the SoA / updateLineX cell the hand-written kernel is bound to: both through the reference's
 * SerialSimulator (VanillaUpdateFunctor-free FixedCoord path vs FixedNeighborhoodUpdateFunctor), member by member. */
class CavityAoSInitializer : public SimpleInitializer<LBMCellAoS>
{
public:
    CavityAoSInitializer(const Coord<3>& dim, unsigned steps) : SimpleInitializer<LBMCellAoS>(dim, steps) {}

    virtual void grid(GridBase<LBMCellAoS, 3> *ret)
    {
        // the very cells LBMInitializer produces, converted member by member
        CoordBox<3> box = ret->boundingBox();
        SoAGrid<LBMCellF, Topologies::Cube<3>::Topology> tmp(box);
        LBMInitializer(gridDimensions(), 1).grid(&tmp);
        for (CoordBox<3>::Iterator i = box.begin(); i != box.end(); ++i) {
            ret->set(*i, convert(tmp.get(*i)));
        }
    }

    static LBMCellAoS convert(const LBMCellF& c)
    {
        LBMCellAoS a(c.C, c.state);
        a.N = c.N; a.E = c.E; a.W = c.W; a.S = c.S; a.T = c.T; a.B = c.B;
        a.NW = c.NW; a.SW = c.SW; a.NE = c.NE; a.SE = c.SE;
        a.TW = c.TW; a.BW = c.BW; a.TE = c.TE; a.BE = c.BE;
        a.TN = c.TN; a.BN = c.BN; a.TS = c.TS; a.BS = c.BS;
        a.density = c.density; a.velocityX = c.velocityX; a.velocityY = c.velocityY; a.velocityZ = c.velocityZ;
        return a;
    }
};

static void testAoSAndSoALBMModelsAgree()
{
    Coord<3> dim(14, 9, 8);
    unsigned steps = 12;
    SerialSimulator<LBMCellF> soa(new LBMInitializer(dim, steps));
    SerialSimulator<LBMCellAoS> aos(new CavityAoSInitializer(dim, steps));
    soa.run();
    aos.run();
    long bad = 0;
    CoordBox<3> box(Coord<3>(), dim);
    for (CoordBox<3>::Iterator i = box.begin(); i != box.end(); ++i) {
        if (!(CavityAoSInitializer::convert(soa.getGrid()->get(*i)) == aos.getGrid()->get(*i))) ++bad;
    }
    CHECK(bad == 0);
    std::printf("LBM: AoS update() model vs SoA updateLineX model through SerialSimulator: %ld differing cells after %u steps\n", bad, steps);
}

/* A writer that is ONLY a ParallelWriter (what programs written for the reference's StripingSimulator / HiParSimulator
 * register, parallelization/distributedsimulator.h:43-46): B200StripingSimulator calls it once per event with the whole
 * simulation area, rank 0, lastCall = true, at the steps a serial Writer of the same period sees on SerialSimulator, and
 * the grid it is handed holds the reference's state of that step. */
struct WriterCall {
    unsigned step;
    int event;
    double sum;
    bool operator==(const WriterCall& o) const { return step == o.step && event == o.event && sum == o.sum; }
};

template<typename GRID>
static double gridSum(const GRID& grid)
{
    double sum = 0;
    CoordBox<3> box = grid.boundingBox();
    for (CoordBox<3>::Iterator i = box.begin(); i != box.end(); ++i) {
        sum += grid.get(*i).temp;
    }
    return sum;
}

class SerialProbe : public Clonable<Writer<Jacobi7Cube>, SerialProbe>
{
public:
    SerialProbe(unsigned period, std::vector<WriterCall> *log) : Clonable<Writer<Jacobi7Cube>, SerialProbe>("", period), log(log) {}

    virtual void stepFinished(const GridType& grid, unsigned step, WriterEvent event)
    {
        WriterCall c = {step, (int)event, gridSum(grid)};
        log->push_back(c);
    }

private:
    std::vector<WriterCall> *log;
};

class ParallelProbe : public Clonable<ParallelWriter<Jacobi7Cube>, ParallelProbe>
{
public:
    ParallelProbe(unsigned period, std::vector<WriterCall> *log, int *bad, int *regionSet) :
        Clonable<ParallelWriter<Jacobi7Cube>, ParallelProbe>("", period), log(log), bad(bad), regionSet(regionSet) {}

    virtual void setRegion(const Region<3>& newRegion)
    {
        ParallelWriter<Jacobi7Cube>::setRegion(newRegion);
        ++*regionSet;
    }

    virtual void stepFinished(const GridType& grid, const RegionType& validRegion, const CoordType& globalDimensions,
                              unsigned step, WriterEvent event, std::size_t rank, bool lastCall)
    {
        *bad += !(validRegion == region) || validRegion.size() != (std::size_t)globalDimensions.prod() || rank != 0 || !lastCall ||
            !(grid.boundingBox().dimensions == globalDimensions);
        WriterCall c = {step, (int)event, gridSum(grid)};
        log->push_back(c);
    }

private:
    std::vector<WriterCall> *log;
    int *bad;
    int *regionSet;
};

static void testParallelWritersOnTheStripingSimulator()
{
    typedef Jacobi7Cube CELL;
    Coord<3> dim(9, 6, 12);
    for (int fuse = 0; fuse < 2; ++fuse) {
        std::vector<WriterCall> a, b;
        int bad = 0, regionSet = 0;
        SerialSimulator<CELL> ref(new SeededInitializer<CELL>(dim, 10));
        ref.addWriter(new SerialProbe(4, &a));
        ref.run();
        B200StripingSimulator<CELL> sim(new SeededInitializer<CELL>(dim, 10), std::vector<int>(3, 0), 2);
        sim.fuseSteps = fuse != 0;
        sim.addWriter(new ParallelProbe(4, &b, &bad, &regionSet));
        sim.addWriter(new TracingWriter<CELL>(100, 10, 0, std::cout));   // both a Writer and a ParallelWriter: the serial overload, no ambiguity
        sim.run();
        CHECK(a.size() == 4 && a == b);   // INITIALIZED @0, STEP_FINISHED @4 and @8, ALL_DONE @10
        CHECK(bad == 0 && regionSet == 1);
        std::printf("ParallelWriter on B200StripingSimulator (fuseSteps %d): %zu calls, %s the serial Writer's on SerialSimulator\n",
                    fuse, b.size(), a == b ? "same steps, events and grids as" : "DIFFERENT from");
    }
}

/* call sites written for StripingSimulator(init, balancer, period) / HiParSimulator(init, balancer, period, ghostZoneWidth) */
class CountingBalancer : public LoadBalancer
{
public:
    explicit CountingBalancer(int *alive) : alive(alive) { ++*alive; }
    virtual ~CountingBalancer() { --*alive; }
    virtual WeightVec balance(const WeightVec& weights, const LoadVec&) { ++asked; return weights; }
    static int asked;

private:
    int *alive;
};
int CountingBalancer::asked = 0;

static void testReferenceConstructorSignatures()
{
    typedef Jacobi7Cube CELL;
    Coord<3> dim(9, 6, 12);
    int alive = 0;
    SerialSimulator<CELL> ref(new SeededInitializer<CELL>(dim, 5));
    ref.run();
    {
        B200StripingSimulator<CELL> striping(new SeededInitializer<CELL>(dim, 5), new CountingBalancer(&alive), 2);
        B200StripingSimulator<CELL> hipar(new SeededInitializer<CELL>(dim, 5), new CountingBalancer(&alive), 1, 2);
        B200StripingSimulator<CELL> none(new SeededInitializer<CELL>(dim, 5), 0);
        CHECK(alive == 2);
        striping.run();
        hipar.run();
        none.run();
        long bad = 0;
        CoordBox<3> box(Coord<3>(), dim);
        for (CoordBox<3>::Iterator i = box.begin(); i != box.end(); ++i) {
            double want = ref.getGrid()->get(*i).temp;
            bad += striping.getGrid()->get(*i).temp != want || hipar.getGrid()->get(*i).temp != want || none.getGrid()->get(*i).temp != want;
        }
        CHECK(bad == 0);
    }
    CHECK(alive == 0 && CountingBalancer::asked == 0);
    {
        // CUDASimulator(initializer, blockSize) and OpenMPSimulator(initializer, enableFineGrainedParallelism) call sites
        B200Simulator<CELL> cuda(new SeededInitializer<CELL>(dim, 5), Coord<3>(128, 4, 1));
        B200Simulator<CELL> omp(new SeededInitializer<CELL>(dim, 5), true);     // a bool is not a device id
        B200Simulator<CELL> plain(new SeededInitializer<CELL>(dim, 5), 0);      // an int is
        B200Simulator<CELL> second(new SeededInitializer<CELL>(dim, 5), 1);
        int devices[4] = {-1, -1, -1, -1};
        B200Simulator<CELL> *sims[4] = {&cuda, &omp, &plain, &second};
        for (int i = 0; i < 4; ++i) {
            const B200Grid<CELL> *g = dynamic_cast<const B200Grid<CELL>*>(sims[i]->getGrid());
            CHECK(g != 0);
            b200geo_grid_device(const_cast<B200Grid<CELL>*>(g)->raw(), &devices[i]);
        }
        CHECK(devices[0] == 0 && devices[1] == 0 && devices[2] == 0 && devices[3] == 1);
        cuda.run();
        omp.run();
        plain.run();
        long bad = 0;
        CoordBox<3> box(Coord<3>(), dim);
        for (CoordBox<3>::Iterator i = box.begin(); i != box.end(); ++i) {
            double want = ref.getGrid()->get(*i).temp;
            bad += cuda.getGrid()->get(*i).temp != want || omp.getGrid()->get(*i).temp != want || plain.getGrid()->get(*i).temp != want;
        }
        CHECK(bad == 0);
    }
    std::printf("reference constructor signatures (initializer, balancer, period[, ghostZoneWidth]): balancers owned and released, results identical\n");
}

int main()
{
    try {
        NBodyParams::dt() = 0.01;
        testReferenceConstructorSignatures();
        testParallelWritersOnTheStripingSimulator();
        testAoSAndSoALBMModelsAgree();
        testCapacityExceeded();
        testResizeAndWriteOrder();
        testSelectorsAndErrors();
        testMemberBoxFastPath();
        testMemberReadAhead();
    } catch (const std::exception& e) {
        std::printf("FAILED with exception: %s\n", e.what());
        return 2;
    }
    if (failures) {
        std::printf("%d check(s) FAILED\n", failures);
        return 1;
    }
    std::printf("host_logic_test: all checks passed\n");
    return 0;
}
