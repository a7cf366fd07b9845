/* Drop-in test of the C++ façade: the SAME user models (oracle/models/*.h), the reference's own
 * Initializer / MockWriter / MockSteerer / Selector / Region classes, and for every case the
 * reference's SerialSimulator running beside B200Simulator in one process. Compiled HERE against
 * /root/reference (tests/facade/Makefile) — the binary travels to the GPU box and is run by
 * tests/test_facade_gpu.py. Exit code 0 = all checks passed. */
#include "fixtures.h"

#include <ctime>

template<typename CELL, typename INIT, int DIM>
void compareWithSerialSimulator(const char *name, const Coord<DIM>& dim, unsigned steps)
{
    SerialSimulator<CELL> ref(new INIT(dim, steps));
    B200Simulator<CELL> sim(new INIT(dim, steps));
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
    std::printf("%-16s %s: %ld differing cells after %u steps\n", name, bad ? "MISMATCH" : "bit-exact", bad, steps);

    // step() advances exactly one step, like the reference
    B200Simulator<CELL> single(new INIT(dim, steps));
    single.step();
    CHECK(single.getStep() == 1);
}

template<typename SIM>
void recordEvents(SIM& sim, std::vector<std::string> *log)
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

void testEventProtocol()
{
    typedef Jacobi7Cube CELL;
    std::vector<std::string> a, b;
    {
        SerialSimulator<CELL> ref(new SeededInitializer<CELL>(Coord<3>(8, 6, 5), 7));
        recordEvents(ref, &a);
    }
    {
        B200Simulator<CELL> sim(new SeededInitializer<CELL>(Coord<3>(8, 6, 5), 7));
        recordEvents(sim, &b);
    }
    CHECK(a.size() > 10);
    CHECK(a == b);
    std::printf("event protocol: %zu writer/steerer events, %s\n", a.size(), a == b ? "identical" : "DIFFERENT");
}

void testRegionBytesMatchSoAGrid()
{
    typedef LBMCellF CELL;
    Coord<3> dim(12, 7, 5);
    CoordBox<3> box(Coord<3>(), dim);
    SoAGrid<CELL, Topologies::Cube<3>::Topology> ref(box);
    B200Grid<CELL> dev(box);
    LBMInitializer init(dim, 1);
    init.grid(&ref);
    init.grid(&dev);

    Region<3> region;
    region << Streak<3>(Coord<3>(0, 0, 0), 12) << Streak<3>(Coord<3>(3, 2, 1), 9)
           << Streak<3>(Coord<3>(11, 6, 4), 12) << Streak<3>(Coord<3>(1, 6, 3), 4);
    std::vector<char> a, b;
    ref.saveRegion(&a, region);
    dev.saveRegion(&b, region);
    CHECK(a.size() == b.size());
    CHECK(a == b);

    // loadRegion round trip: write the reference's bytes shifted by one cell in x, read back
    Region<3> target;
    target << Streak<3>(Coord<3>(2, 3, 2), 8);
    Region<3> source;
    source << Streak<3>(Coord<3>(3, 2, 1), 9);
    std::vector<char> chunk;
    ref.saveRegion(&chunk, source);
    dev.loadRegion(chunk, target);
    ref.loadRegion(chunk, target);
    std::vector<char> c, d;
    ref.saveRegion(&c, target);
    dev.saveRegion(&d, target);
    CHECK(c == d);
    CHECK(ref.get(Coord<3>(4, 3, 2)) == dev.get(Coord<3>(4, 3, 2)));

    bool thrown = false;
    try {
        std::vector<char> tooShort(3);
        dev.loadRegion(tooShort, target);
    } catch (const std::invalid_argument&) {
        thrown = true;
    }
    CHECK(thrown);

    // Selector-based member I/O (GridBase::saveMember / loadMember, storage/gridbase.h:217-261)
    Selector<CELL> sel(&CELL::density, "density");
    std::vector<float> da(region.size()), db(region.size());
    ref.saveMember(da.data(), MemoryLocation::HOST, sel, region);
    dev.saveMember(db.data(), MemoryLocation::HOST, sel, region);
    CHECK(da == db);
    for (std::size_t i = 0; i < da.size(); ++i) da[i] = 3.0f + i;
    ref.loadMember(da.data(), MemoryLocation::HOST, sel, region);
    dev.loadMember(da.data(), MemoryLocation::HOST, sel, region);
    CHECK(ref.get(Coord<3>(5, 2, 1)) == dev.get(Coord<3>(5, 2, 1)));
    bool wrongType = false;
    try {
        std::vector<double> wrong(region.size());
        dev.saveMember(wrong.data(), MemoryLocation::HOST, sel, region);
    } catch (const std::invalid_argument&) {
        wrongType = true;
    }
    CHECK(wrongType);
    std::printf("saveRegion/loadRegion/saveMember/loadMember vs SoAGrid: %s\n", (a == b && c == d) ? "byte-identical" : "DIFFERENT");
}

template<typename REAL>
static void compareNBody(const char *name, const Coord<3>& dim, unsigned steps)
{
    typedef BoxCell<FixedArray<LJParticle<REAL>, NBODY_CAPACITY> > Cell;
    NBodyParams::dt() = 0.01;
    SerialSimulator<Cell> ref(new ParticleInitializer<REAL>(dim, steps));
    B200Simulator<Cell> dev(new ParticleInitializer<REAL>(dim, steps));
    ref.run();
    dev.run();
    const GridBase<Cell, 3> *a = ref.getGrid();
    const GridBase<Cell, 3> *b = dev.getGrid();
    std::size_t particles = 0, moved = 0, bad = 0;
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
            moved += ca[p].vel[0] != 0;
        }
    }
    CHECK(bad == 0);
    CHECK(particles > 0 && moved > 0);
    CHECK(ref.getStep() == dev.getStep());
    std::printf("%-14s %3dx%3dx%3d x %2u steps: %zu particles, %s\n", name, dim.x(), dim.y(), dim.z(), steps, particles,
                bad == 0 ? "bit-identical to SerialSimulator" : "DIFFERENT");
}

/* Initializers / Writers that go cell by cell (GridBase::set(Coord) / get(Coord) in a loop, as the reference's
 * examples do): writes are combined and rows cached, later writes win, a read after a write sees the write */
static void testCellByCellAccess()
{
    typedef Jacobi7Cube CELL;
    Coord<3> dim(96, 64, 48);
    CoordBox<3> box(Coord<3>(), dim);
    B200Grid<CELL> dev(box);
    std::clock_t t0 = std::clock();
    for (CoordBox<3>::Iterator i = box.begin(); i != box.end(); ++i) {
        dev.set(*i, CELL(uniform(i->toIndex(dim))));
    }
    // overwrite a few cells twice in a row, then read them back at once (write, write, read)
    dev.set(Coord<3>(5, 6, 7), CELL(1.25));
    dev.set(Coord<3>(5, 6, 7), CELL(2.5));
    dev.set(Streak<3>(Coord<3>(4, 6, 7), 7), std::vector<CELL>(3, CELL(7.0)).data());
    dev.set(Coord<3>(6, 6, 7), CELL(9.0));
    CHECK(dev.get(Coord<3>(4, 6, 7)) == CELL(7.0));
    CHECK(dev.get(Coord<3>(5, 6, 7)) == CELL(7.0));
    CHECK(dev.get(Coord<3>(6, 6, 7)) == CELL(9.0));
    dev.set(Coord<3>(6, 6, 7), CELL(uniform(Coord<3>(6, 6, 7).toIndex(dim))));
    dev.set(Coord<3>(5, 6, 7), CELL(uniform(Coord<3>(5, 6, 7).toIndex(dim))));
    dev.set(Coord<3>(4, 6, 7), CELL(uniform(Coord<3>(4, 6, 7).toIndex(dim))));
    long bad = 0;
    for (CoordBox<3>::Iterator i = box.begin(); i != box.end(); ++i) {
        if (!(dev.get(*i) == CELL(uniform(i->toIndex(dim))))) ++bad;
    }
    double seconds = (double)(std::clock() - t0) / CLOCKS_PER_SEC;
    CHECK(bad == 0);
    // the edge ring is read straight from the device
    dev.setEdge(CELL(0.125));
    CHECK(dev.get(Coord<3>(-1, 0, 0)) == CELL(0.125));
    // a sweep invalidates cached rows
    CELL before = dev.get(Coord<3>(10, 10, 10));
    dev.update(0, 1);
    CHECK(!(dev.get(Coord<3>(10, 10, 10)) == before));
    std::printf("cell-by-cell set + get of %d cells: %ld wrong, %.2f s CPU\n", dim.prod(), bad, seconds);
}

int main()
{
    try {
        compareWithSerialSimulator<Jacobi7Cube, SeededInitializer<Jacobi7Cube>, 3>("Jacobi7Cube", Coord<3>(34, 9, 11), 9);
        compareWithSerialSimulator<Jacobi7Torus, SeededInitializer<Jacobi7Torus>, 3>("Jacobi7Torus", Coord<3>(16, 5, 7), 6);
        compareWithSerialSimulator<Jacobi27Cube, SeededInitializer<Jacobi27Cube>, 3>("Jacobi27Cube", Coord<3>(21, 10, 8), 7);
        compareWithSerialSimulator<Jacobi27Torus, SeededInitializer<Jacobi27Torus>, 3>("Jacobi27Torus", Coord<3>(12, 6, 6), 5);
        compareWithSerialSimulator<Jacobi6Cube, SeededInitializer<Jacobi6Cube>, 3>("Jacobi6Cube", Coord<3>(10, 12, 9), 8);
        compareWithSerialSimulator<Jacobi6Torus, SeededInitializer<Jacobi6Torus>, 3>("Jacobi6Torus", Coord<3>(20, 6, 4), 8);
        compareWithSerialSimulator<ConwayCube, SeededInitializer<ConwayCube>, 2>("ConwayCube", Coord<2>(70, 33), 25);
        compareWithSerialSimulator<ConwayTorus, SeededInitializer<ConwayTorus>, 2>("ConwayTorus", Coord<2>(48, 20), 25);
        compareWithSerialSimulator<LBMCellF, LBMInitializer, 3>("LBMCellF", Coord<3>(18, 12, 10), 15);
        compareNBody<float>("NBody<float>", Coord<3>(7, 5, 4), 8);
        compareNBody<double>("NBody<double>", Coord<3>(4, 6, 3), 6);
        testEventProtocol();
        testRegionBytesMatchSoAGrid();
        testCellByCellAccess();
        {
            B200Simulator<Jacobi7Cube> sim(new SeededInitializer<Jacobi7Cube>(Coord<3>(20, 7, 12), 6));
            testSerialBOVWriter(sim, "single");
        }
    } catch (const std::exception& e) {
        std::printf("FAILED with exception: %s\n", e.what());
        return 2;
    }
    if (failures) {
        std::printf("%d check(s) FAILED\n", failures);
        return 1;
    }
    std::printf("facade_test: all checks passed\n");
    return 0;
}
