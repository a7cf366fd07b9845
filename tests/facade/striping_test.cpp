/* Drop-in test of the single-process multi-GPU façade (include/libgeodecomp_b200/b200stripingsimulator.h):
 * the user models of oracle/models/*.h on 1..4 slabs, spread round-robin over the GPUs present (on a
 * 1-GPU box all slabs share device 0 — same schedule, the halo copies are device-local), the reference's
 * SerialSimulator beside it in this process. Required: bit-identical grids for every slab count and
 * ghost-zone width, the event protocol of SerialSimulator with the reference's MockWriter / MockSteerer,
 * and SoAGrid-compatible region / member byte streams across slab boundaries.
 * Compiled HERE (tests/facade/Makefile), run on the GPU box by tests/test_facade_gpu.py. */
#include "fixtures.h"

#include <libgeodecomp_b200/b200stripingsimulator.h>

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
void compareStriped(const char *name, const Coord<DIM>& dim, unsigned steps, int slabs, int ghost)
{
    SerialSimulator<CELL> ref(new INIT(dim, steps));
    B200StripingSimulator<CELL> sim(new INIT(dim, steps), devicesFor(slabs), ghost);
    ref.run();
    sim.run();
    CHECK(ref.getStep() == steps);
    CHECK(sim.getStep() == steps);
    const GridBase<CELL, DIM> *a = ref.getGrid();
    const GridBase<CELL, DIM> *b = sim.getGrid();
    CHECK(a->boundingBox() == b->boundingBox());
    CHECK(a->getEdge() == b->getEdge());
    long bad = 0;
    CoordBox<DIM> box = a->boundingBox();
    std::vector<CELL> ra(dim.x()), rb(dim.x());
    for (typename CoordBox<DIM>::StreakIterator i = box.beginStreak(); i != box.endStreak(); ++i) {
        a->get(*i, ra.data());
        b->get(*i, rb.data());
        for (int x = 0; x < dim.x(); ++x) {
            if (!(ra[x] == rb[x])) ++bad;
        }
    }
    CHECK(bad == 0);
    std::pair<unsigned long long, unsigned long long> st = sim.stripedGrid().exchangeStatistics();
    CHECK(slabs == 1 || st.first > 0);
    std::printf("%-14s %d slab(s), ghost width %d: %s (%ld differing cells after %u steps; %llu exchanges, %.2f MB shipped)\n",
                name, slabs, ghost, bad ? "MISMATCH" : "bit-exact", bad, steps, st.first, st.second / 1e6);
}

template<typename SIM>
static void recordEvents(SIM& sim, std::vector<std::string> *log)
{
    typedef Jacobi7Cube CELL;
    typedef MockWriter<CELL>::EventsStore WEvents;
    typedef MockSteerer<CELL>::EventsStore SEvents;
    SharedPtr<WEvents>::Type w1(new WEvents), w3(new WEvents);
    SharedPtr<SEvents>::Type s2(new SEvents);
    sim.addWriter(new MockWriter<CELL>(w1, 1));
    sim.addWriter(new MockWriter<CELL>(w3, 3));
    sim.addSteerer(new MockSteerer<CELL>(2, s2));
    sim.run();
    for (std::size_t i = 0; i < w1->size(); ++i) log->push_back("w1 " + (*w1)[i].toString());
    for (std::size_t i = 0; i < w3->size(); ++i) log->push_back("w3 " + (*w3)[i].toString());
    for (std::size_t i = 0; i < s2->size(); ++i) log->push_back("s2 " + (*s2)[i].toString());
}

static void testEventProtocol()
{
    typedef Jacobi7Cube CELL;
    std::vector<std::string> a, b;
    {
        SerialSimulator<CELL> ref(new SeededInitializer<CELL>(Coord<3>(8, 6, 12), 7));
        recordEvents(ref, &a);
    }
    {
        B200StripingSimulator<CELL> sim(new SeededInitializer<CELL>(Coord<3>(8, 6, 12), 7), devicesFor(3), 2);
        recordEvents(sim, &b);
    }
    CHECK(a.size() > 10);
    CHECK(a == b);
    std::printf("event protocol on 3 slabs: %zu writer/steerer events, %s\n", a.size(), a == b ? "identical" : "DIFFERENT");
}

/* a Steerer that overwrites a plane next to a slab boundary mid-run: the neighbour's ghost copy must be
 * refreshed before the next sweep (GridBase::set from a plugin, io/steerer.h:113-121) */
class PokeSteerer : public Steerer<Jacobi7Cube>
{
public:
    typedef Steerer<Jacobi7Cube>::SteererFeedback SteererFeedback;
    typedef Steerer<Jacobi7Cube>::GridType GridType;

    explicit PokeSteerer(int z) : Steerer<Jacobi7Cube>(3), z(z) {}

    virtual void nextStep(GridType *grid, const Region<3>&, const Coord<3>& dims, unsigned step, SteererEvent event,
                          std::size_t, bool, SteererFeedback *)
    {
        if (event != STEERER_NEXT_STEP || step == 0) {
            return;
        }
        std::vector<Jacobi7Cube> row(dims.x(), Jacobi7Cube(2.0 + step));
        for (int y = 0; y < dims.y(); ++y) {
            grid->set(Streak<3>(Coord<3>(0, y, z), dims.x()), row.data());
        }
    }

private:
    int z;
};

static void testSteererWritesAcrossSlabs()
{
    typedef Jacobi7Cube CELL;
    Coord<3> dim(10, 7, 12);
    SerialSimulator<CELL> ref(new SeededInitializer<CELL>(dim, 8));
    B200StripingSimulator<CELL> sim(new SeededInitializer<CELL>(dim, 8), devicesFor(2), 2);
    ref.addSteerer(new PokeSteerer(5));     // last plane of slab 0 = ghost plane of slab 1
    sim.addSteerer(new PokeSteerer(5));
    ref.run();
    sim.run();
    long bad = 0;
    CoordBox<3> box(Coord<3>(), dim);
    for (CoordBox<3>::Iterator i = box.begin(); i != box.end(); ++i) {
        if (!(ref.getGrid()->get(*i) == sim.getGrid()->get(*i))) ++bad;
    }
    CHECK(bad == 0);
    std::printf("steerer writing next to a slab boundary: %s\n", bad ? "MISMATCH" : "bit-exact");
}

static void testRegionAndMemberStreamsAcrossSlabs()
{
    typedef LBMCellF CELL;
    Coord<3> dim(12, 7, 9);
    CoordBox<3> box(Coord<3>(), dim);
    SoAGrid<CELL, Topologies::Cube<3>::Topology> ref(box);
    B200StripedGrid<CELL> dev(box, devicesFor(3), 1);
    LBMInitializer init(dim, 1);
    init.grid(&ref);
    init.grid(&dev);

    Region<3> region;   // streaks in slabs 0, 1 and 2 (planes 0-2, 3-5, 6-8)
    region << Streak<3>(Coord<3>(0, 0, 0), 12) << Streak<3>(Coord<3>(3, 2, 2), 9) << Streak<3>(Coord<3>(1, 6, 3), 4)
           << Streak<3>(Coord<3>(2, 1, 5), 11) << Streak<3>(Coord<3>(11, 6, 8), 12) << Streak<3>(Coord<3>(0, 0, 7), 5);
    std::vector<char> a, b;
    ref.saveRegion(&a, region);
    dev.saveRegion(&b, region);
    CHECK(a.size() == b.size());
    CHECK(a == b);

    Region<3> source, target;
    source << Streak<3>(Coord<3>(3, 2, 2), 9) << Streak<3>(Coord<3>(2, 1, 5), 8);
    target << Streak<3>(Coord<3>(0, 3, 4), 6) << Streak<3>(Coord<3>(4, 3, 6), 10);
    std::vector<char> chunk;
    ref.saveRegion(&chunk, source);
    ref.loadRegion(chunk, target);
    dev.loadRegion(chunk, target);
    std::vector<char> c, d;
    ref.saveRegion(&c, target);
    dev.saveRegion(&d, target);
    CHECK(c == d);
    CHECK(ref.get(Coord<3>(5, 3, 6)) == dev.get(Coord<3>(5, 3, 6)));

    Selector<CELL> sel(&CELL::density, "density");
    std::vector<float> da(region.size()), db(region.size());
    ref.saveMember(da.data(), MemoryLocation::HOST, sel, region);
    dev.saveMember(db.data(), MemoryLocation::HOST, sel, region);
    CHECK(da == db);
    for (std::size_t i = 0; i < da.size(); ++i) da[i] = 3.0f + i;
    ref.loadMember(da.data(), MemoryLocation::HOST, sel, region);
    dev.loadMember(da.data(), MemoryLocation::HOST, sel, region);
    CHECK(ref.get(Coord<3>(3, 1, 5)) == dev.get(Coord<3>(3, 1, 5)));
    CHECK(ref.get(Coord<3>(2, 0, 7)) == dev.get(Coord<3>(2, 0, 7)));
    std::printf("saveRegion/loadRegion/saveMember/loadMember across 3 slabs vs SoAGrid: %s\n",
                (a == b && c == d && da.size() == db.size()) ? "byte-identical" : "DIFFERENT");
}

template<typename REAL>
static void compareNBodyStriped(const char *name, const Coord<3>& dim, unsigned steps, int slabs)
{
    typedef BoxCell<FixedArray<LJParticle<REAL>, NBODY_CAPACITY> > Cell;
    NBodyParams::dt() = 0.01;
    SerialSimulator<Cell> ref(new ParticleInitializer<REAL>(dim, steps));
    B200StripingSimulator<Cell> dev(new ParticleInitializer<REAL>(dim, steps), devicesFor(slabs));
    ref.run();
    dev.run();
    const GridBase<Cell, 3> *a = ref.getGrid();
    const GridBase<Cell, 3> *b = dev.getGrid();
    std::size_t particles = 0, bad = 0;
    CoordBox<3> box(Coord<3>(), dim);
    for (CoordBox<3>::Iterator i = box.begin(); i != box.end(); ++i) {
        Cell ca = a->get(*i), cb = b->get(*i);
        if (ca.size() != cb.size()) {
            ++bad;
            continue;
        }
        particles += ca.size();
        for (std::size_t p = 0; p < ca.size(); ++p) {
            if (std::memcmp(ca[p].pos, cb[p].pos, sizeof(ca[p].pos)) || std::memcmp(ca[p].vel, cb[p].vel, sizeof(ca[p].vel))) {
                ++bad;
            }
        }
    }
    CHECK(bad == 0);
    CHECK(particles > 0);
    CHECK(ref.getStep() == dev.getStep());
    std::pair<unsigned long long, unsigned long long> st = dev.stripedGrid().exchangeStatistics();
    CHECK(slabs == 1 || st.first >= steps);
    std::printf("%-14s %d slab(s) of containers: %zu particles after %u steps, %s (%llu exchanges)\n", name, slabs, particles, steps,
                bad == 0 ? "bit-identical to SerialSimulator" : "DIFFERENT", st.first);
}

int main()
{
    try {
        int devices = b200geo_device_count();
        std::printf("striping_test: %d CUDA device(s)\n", devices);
        compareStriped<Jacobi7Cube, SeededInitializer<Jacobi7Cube>, 3>("Jacobi7Cube", Coord<3>(34, 9, 22), 9, 1, 1);
        compareStriped<Jacobi7Cube, SeededInitializer<Jacobi7Cube>, 3>("Jacobi7Cube", Coord<3>(34, 9, 22), 9, 2, 1);
        compareStriped<Jacobi7Cube, SeededInitializer<Jacobi7Cube>, 3>("Jacobi7Cube", Coord<3>(34, 9, 22), 9, 3, 2);
        compareStriped<Jacobi27Cube, SeededInitializer<Jacobi27Cube>, 3>("Jacobi27Cube", Coord<3>(21, 10, 32), 11, 4, 4);
        compareStriped<Jacobi27Torus, SeededInitializer<Jacobi27Torus>, 3>("Jacobi27Torus", Coord<3>(12, 6, 18), 7, 2, 2);
        compareStriped<Jacobi27Torus, SeededInitializer<Jacobi27Torus>, 3>("Jacobi27Torus", Coord<3>(12, 6, 18), 7, 3, 3);
        compareStriped<Jacobi6Torus, SeededInitializer<Jacobi6Torus>, 3>("Jacobi6Torus", Coord<3>(20, 6, 9), 8, 3, 1);
        compareStriped<ConwayCube, SeededInitializer<ConwayCube>, 2>("ConwayCube", Coord<2>(70, 33), 25, 3, 1);
        compareStriped<ConwayTorus, SeededInitializer<ConwayTorus>, 2>("ConwayTorus", Coord<2>(64, 20), 25, 2, 2);
        compareStriped<LBMCellF, LBMInitializer, 3>("LBMCellF", Coord<3>(18, 12, 10), 15, 2, 1);
        compareStriped<LBMCellF, LBMInitializer, 3>("LBMCellF", Coord<3>(18, 12, 13), 9, 3, 2);
        compareNBodyStriped<float>("NBody<float>", Coord<3>(7, 5, 8), 8, 3);
        compareNBodyStriped<float>("NBody<float>", Coord<3>(5, 4, 4), 6, 4);
        compareNBodyStriped<double>("NBody<double>", Coord<3>(4, 6, 5), 6, 2);
        compareNBodyStriped<double>("NBody<double>", Coord<3>(4, 3, 3), 4, 1);
        testEventProtocol();
        testSteererWritesAcrossSlabs();
        testRegionAndMemberStreamsAcrossSlabs();
        {
            B200StripingSimulator<Jacobi7Cube> sim(new SeededInitializer<Jacobi7Cube>(Coord<3>(20, 7, 12), 6), devicesFor(3), 2);
            testSerialBOVWriter(sim, "striped");
        }
        bool thrown = false;
        try {
            B200StripingSimulator<Jacobi7Cube> tooThin(new SeededInitializer<Jacobi7Cube>(Coord<3>(8, 8, 4), 1), devicesFor(4), 2);
        } catch (const std::invalid_argument&) {
            thrown = true;
        }
        CHECK(thrown);
    } catch (const std::exception& e) {
        std::printf("FAILED with exception: %s\n", e.what());
        return 2;
    }
    if (failures) {
        std::printf("%d check(s) FAILED\n", failures);
        return 1;
    }
    std::printf("striping_test: all checks passed\n");
    return 0;
}
