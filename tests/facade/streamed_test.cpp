/* The streamed run of the C++ façade (include/libgeodecomp_b200/b200streamedrun.h) beside the reference's
 * SerialSimulator: B200Simulator::run() with nothing but ParallelWriters registered pipelines Initializer, sweeps and
 * writers chunk by chunk — and must leave the same grid, bit for bit, hand the writers every cell exactly once per
 * event, and fall back to the plain schedule whenever a plugin needs the whole grid at once.
 *
 * Built twice (tests/facade/Makefile): against libb200geo.so (GPU box) and against mock_b200geo.cpp (CPU suite). */
#include "fixtures.h"

#include <libgeodecomp/io/parallelwriter.h>

#include <map>

/* collects what the simulator hands it: every cell of every validRegion, read the way `mode` says */
template<typename CELL>
class CollectWriter : public ParallelWriter<CELL>
{
public:
    typedef typename ParallelWriter<CELL>::GridType GridType;
    typedef typename ParallelWriter<CELL>::RegionType RegionType;
    typedef typename ParallelWriter<CELL>::CoordType CoordType;
    static const int DIM = APITraits::SelectTopology<CELL>::Value::DIM;

    struct Log {
        std::map<std::pair<unsigned, int>, Region<DIM> > seen;   /* (step, event) -> union of the valid regions */
        std::map<std::pair<unsigned, int>, int> calls, lastCalls, overlaps;
        std::vector<CELL> cells;                                  /* the grid of the last WRITER_ALL_DONE, [z][y][x] */
        Coord<DIM> dim;
    };

    CollectWriter(Log *log, unsigned period, bool streaks) : ParallelWriter<CELL>("", period), log(log), streaks(streaks) {}

    virtual ParallelWriter<CELL> *clone() const
    {
        return new CollectWriter(*this);
    }

    virtual void stepFinished(const GridType& grid, const RegionType& validRegion, const CoordType& globalDimensions, unsigned step,
                              WriterEvent event, std::size_t rank, bool lastCall)
    {
        std::pair<unsigned, int> key(step, (int)event);
        Region<DIM> both = log->seen[key] & validRegion;
        log->overlaps[key] += both.empty() ? 0 : 1;
        log->seen[key] += validRegion;
        ++log->calls[key];
        log->lastCalls[key] += lastCall ? 1 : 0;
        CHECK(rank == 0);
        CHECK(this->region.boundingBox().dimensions == globalDimensions);
        if (event != WRITER_ALL_DONE) {
            return;
        }
        log->dim = globalDimensions;
        log->cells.resize(globalDimensions.prod());
        for (typename Region<DIM>::StreakIterator i = validRegion.beginStreak(); i != validRegion.endStreak(); ++i) {
            CELL *row = &log->cells[i->origin.toIndex(globalDimensions)];
            if (streaks) {
                grid.get(*i, row);
            } else {
                for (int x = i->origin.x(); x < i->endX; ++x) {
                    Coord<DIM> c = i->origin;
                    c.x() = x;
                    row[x - i->origin.x()] = grid.get(c);
                }
            }
        }
    }

private:
    Log *log;
    bool streaks;
};

template<typename CELL, int DIM>
static int differingCells(const GridBase<CELL, DIM>& a, const std::vector<CELL>& cells, const Coord<DIM>& dim)
{
    int bad = 0;
    CoordBox<DIM> box(Coord<DIM>(), dim);
    for (typename CoordBox<DIM>::Iterator i = box.begin(); i != box.end(); ++i) {
        CELL x = a.get(*i);
        bad += std::memcmp(&x, &cells[i->toIndex(dim)], sizeof(CELL)) != 0;
    }
    return bad;
}

template<typename CELL, typename INIT, int DIM>
static void compareStreamed(const char *name, const Coord<DIM>& dim, unsigned steps, int chunks, unsigned period, bool expectStreamed,
                            bool streaks = true)
{
    SerialSimulator<CELL> ref(new INIT(dim, steps));
    ref.run();
    B200Simulator<CELL> sim(new INIT(dim, steps));
    sim.streamChunks = chunks;
    typename CollectWriter<CELL>::Log log;
    sim.addWriter(new CollectWriter<CELL>(&log, period, streaks));
    sim.run();
    CHECK(sim.getStep() == steps);
    CHECK((sim.streamedRuns() == 1) == expectStreamed);
    /* the writer was handed the final grid, every cell exactly once, lastCall exactly once per event */
    Region<DIM> whole;
    whole << CoordBox<DIM>(Coord<DIM>(), dim);
    std::pair<unsigned, int> done(steps, (int)WRITER_ALL_DONE), init(0u, (int)WRITER_INITIALIZED);
    CHECK(log.seen[done] == whole);
    CHECK(log.seen[init] == whole);
    CHECK(log.overlaps[done] == 0 && log.overlaps[init] == 0);
    CHECK(log.lastCalls[done] == 1 && log.lastCalls[init] == 1);
    CHECK(!expectStreamed || log.calls[done] > 1);
    int bad = differingCells(*ref.getGrid(), log.cells, dim);
    CHECK(bad == 0);
    /* ... and the simulator's own grid is that grid, too */
    const GridBase<CELL, DIM> *mine = sim.getGrid();
    std::vector<CELL> pulled(dim.prod());
    CoordBox<DIM> box(Coord<DIM>(), dim);
    for (typename CoordBox<DIM>::StreakIterator i = box.beginStreak(); i != box.endStreak(); ++i) {
        mine->get(*i, &pulled[i->origin.toIndex(dim)]);
    }
    int bad2 = differingCells(*ref.getGrid(), pulled, dim);
    CHECK(bad2 == 0);
    std::printf("%s: %u steps, %d chunks asked for, %s schedule, writer called %d times at the end: %d / %d cells differ from SerialSimulator\n",
                name, steps, chunks, sim.streamedRuns() ? "streamed" : "plain", log.calls[done], bad, bad2);
}

/* whole boxes through Selector I/O — the fast way: Initializer::grid loads its bounding box with loadMember, the
 * writer pulls its validRegion with saveMember (what tests/facade/e2e_bench.cpp times at 1024^3) */
class MemberInitializer : public SimpleInitializer<Jacobi27Cube>
{
public:
    MemberInitializer(const Coord<3>& dim, unsigned steps) : SimpleInitializer<Jacobi27Cube>(dim, steps), field(dim.prod())
    {
        for (std::size_t i = 0; i < field.size(); ++i) {
            field[i] = uniform(i);
        }
    }

    virtual void grid(GridBase<Jacobi27Cube, 3> *ret)
    {
        CoordBox<3> box = ret->boundingBox();
        ret->setEdge(Jacobi27Cube(0.25));
        Region<3> region;
        region << box;
        /* whole planes: contiguous in the host array */
        ret->loadMember(&field[(std::size_t)box.origin.z() * gridDimensions().y() * gridDimensions().x()], MemoryLocation::HOST,
                        Selector<Jacobi27Cube>(&Jacobi27Cube::temp, "temp"), region);
    }

private:
    std::vector<double> field;
};

class MemberWriter : public ParallelWriter<Jacobi27Cube>
{
public:
    MemberWriter(std::vector<double> *out, int *calls) : ParallelWriter<Jacobi27Cube>("", 1u << 30), out(out), calls(calls) {}

    virtual ParallelWriter<Jacobi27Cube> *clone() const
    {
        return new MemberWriter(*this);
    }

    virtual void stepFinished(const GridType& grid, const RegionType& validRegion, const CoordType& globalDimensions, unsigned,
                              WriterEvent event, std::size_t, bool)
    {
        if (event != WRITER_ALL_DONE) {
            return;
        }
        ++*calls;
        out->resize(globalDimensions.prod());
        CoordBox<3> box = validRegion.boundingBox();
        grid.saveMember(&(*out)[(std::size_t)box.origin.z() * globalDimensions.y() * globalDimensions.x()], MemoryLocation::HOST,
                        Selector<Jacobi27Cube>(&Jacobi27Cube::temp, "temp"), validRegion);
    }

private:
    std::vector<double> *out;
    int *calls;
};

static void testMemberIo()
{
    const Coord<3> dim(24, 10, 48);
    const unsigned steps = 7;
    SerialSimulator<Jacobi27Cube> ref(new MemberInitializer(dim, steps));
    ref.run();
    B200Simulator<Jacobi27Cube> sim(new MemberInitializer(dim, steps));
    sim.streamChunks = 6;
    std::vector<double> out;
    int calls = 0;
    sim.addWriter(new MemberWriter(&out, &calls));
    sim.run();   /* the writer's buffer is complete when run() returns: read it right away */
    CHECK(sim.streamedRuns() == 1);
    CHECK(calls > 1);
    int bad = 0;
    CoordBox<3> box(Coord<3>(), dim);
    for (CoordBox<3>::Iterator i = box.begin(); i != box.end(); ++i) {
        double want = ref.getGrid()->get(*i).temp;
        bad += std::memcmp(&want, &out[i->toIndex(dim)], sizeof(double)) != 0;
    }
    CHECK(bad == 0);
    /* a second run() of the same simulator starts from the Initializer again */
    std::fill(out.begin(), out.end(), -1.0);
    sim.run();
    CHECK(sim.streamedRuns() == 2);
    int again = 0;
    for (CoordBox<3>::Iterator i = box.begin(); i != box.end(); ++i) {
        double want = ref.getGrid()->get(*i).temp;
        again += std::memcmp(&want, &out[i->toIndex(dim)], sizeof(double)) != 0;
    }
    CHECK(again == 0);
    std::printf("Selector I/O by boxes: writer called %d times per run, %d / %d cells differ (first / second run)\n", calls / 2, bad, again);
}

/* a serial Writer wants the whole grid at once: the plain schedule, ParallelWriters served with the whole area */
static void testFallbacks()
{
    typedef Jacobi7Cube CELL;
    const Coord<3> dim(16, 9, 40);
    SerialSimulator<CELL> ref(new SeededInitializer<CELL>(dim, 5));
    ref.run();
    {
        B200Simulator<CELL> sim(new SeededInitializer<CELL>(dim, 5));
        CollectWriter<CELL>::Log log;
        sim.addWriter(new CollectWriter<CELL>(&log, 1000, true));
        SharedPtr<MockWriter<CELL>::EventsStore>::Type events(new MockWriter<CELL>::EventsStore);
        sim.addWriter(new MockWriter<CELL>(events, 5));
        sim.run();
        CHECK(sim.streamedRuns() == 0);
        CHECK(differingCells(*ref.getGrid(), log.cells, dim) == 0);
        CHECK(log.calls[std::make_pair(5u, (int)WRITER_ALL_DONE)] == 1);
    }
    {
        B200Simulator<CELL> sim(new SeededInitializer<CELL>(dim, 5));
        sim.streamIO = false;
        CollectWriter<CELL>::Log log;
        sim.addWriter(new CollectWriter<CELL>(&log, 1000, true));
        sim.run();
        CHECK(sim.streamedRuns() == 0);
        CHECK(differingCells(*ref.getGrid(), log.cells, dim) == 0);
    }
    {
        /* a grid too short to be cut into two chunks */
        B200Simulator<CELL> sim(new SeededInitializer<CELL>(Coord<3>(16, 9, 12), 5));
        CollectWriter<CELL>::Log log;
        sim.addWriter(new CollectWriter<CELL>(&log, 1000, true));
        sim.run();
        CHECK(sim.streamedRuns() == 0);
        CHECK(sim.getStep() == 5);
    }
    std::printf("fallbacks to the plain schedule: serial Writer, streamIO = false, short grid\n");
}

int main()
{
    try {
        /* levels of 2 (27-point), remainder level first */
        compareStreamed<Jacobi27Cube, SeededInitializer<Jacobi27Cube> >("Jacobi27Cube", Coord<3>(20, 12, 40), 7, 8, 1000, true);
        compareStreamed<Jacobi27Cube, SeededInitializer<Jacobi27Cube> >("Jacobi27Cube cell by cell", Coord<3>(9, 5, 33), 4, 4, 1000, true, false);
        /* levels of 4 (7- and 6-point): 9 = 1 + 4 + 4 */
        compareStreamed<Jacobi7Cube, SeededInitializer<Jacobi7Cube> >("Jacobi7Cube", Coord<3>(18, 7, 64), 9, 4, 1000, true);
        compareStreamed<Jacobi6Cube, SeededInitializer<Jacobi6Cube> >("Jacobi6Cube", Coord<3>(5, 3, 50), 3, 5, 1000, true);
        /* LBM: two fused sweeps per level, density / velocity stored by the last level only, state in both buffers */
        compareStreamed<LBMCellF, LBMInitializer>("LBMCellF", Coord<3>(12, 10, 36), 5, 6, 1000, true);
        /* 2-D, one sweep per level: rows are streamed */
        compareStreamed<ConwayCube, SeededInitializer<ConwayCube> >("ConwayCube", Coord<2>(40, 60), 6, 6, 1000, true);
        /* a writer period that falls due inside the run, a Torus: the plain schedule */
        compareStreamed<Jacobi27Cube, SeededInitializer<Jacobi27Cube> >("Jacobi27Cube, period 3", Coord<3>(20, 12, 40), 7, 8, 3, false);
        compareStreamed<Jacobi7Torus, SeededInitializer<Jacobi7Torus> >("Jacobi7Torus", Coord<3>(16, 8, 40), 5, 8, 1000, false);
        testMemberIo();
        testFallbacks();
    } catch (const std::exception& e) {
        std::printf("FAILED with exception: %s\n", e.what());
        return 1;
    }
    if (failures) {
        std::printf("%d checks FAILED\n", failures);
        return 1;
    }
    std::printf("all checks passed\n");
    return 0;
}
