/* Drop-in test of the ContainerCell path of the C++ façade: an ID-keyed model (the mesh element of
 * oracle/models/container.h, i.e. the cell of src/examples/voronoi/main.cpp) in ContainerCell<MeshElement, 16> containers,
 * through the reference's SerialSimulator and through B200Simulator in one process — same Initializer, same
 * cells, temperatures compared bit for bit. Also: GridBase::set/get on B200ContainerGrid (members the bound update does
 * not touch survive, ids ascend, a steering write between steps), the edge container, std::logic_error for an id
 * that is not in the neighbourhood (storage/neighborhoodadapter.h:63-64). Built twice by tests/facade/Makefile: against
 * libb200geo.so (run on the GPU box by tests/test_facade_gpu.py) and against the mock engine (CPU suite). */
#include <libgeodecomp/io/simpleinitializer.h>
#include <libgeodecomp/io/steerer.h>
#include <libgeodecomp/io/writer.h>
#include <libgeodecomp/misc/clonable.h>
#include <libgeodecomp/parallelization/serialsimulator.h>

#include <libgeodecomp_b200/b200simulator.h>

#include "models/container.h"

#include <cstdio>
#include <cstring>

using namespace LibGeoDecomp;
using namespace b200models;

#define BIND(DIM, TORUS) B200GEO_BIND_CARGO(b200models::MeshElement<DIM COMMA TORUS>, temperature, influx, neighborIDs)
#define COMMA ,
BIND(2, false)
BIND(2, true)
BIND(3, false)
BIND(3, true)

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

template<int DIM> struct Wrap {
    static Coord<DIM> torus(Coord<DIM> c, const Coord<DIM>& dim)
    {
        for (int i = 0; i < DIM; ++i) c[i] = (c[i] + dim[i]) % dim[i];
        return c;
    }
};

/* A seeded mesh: container c holds (hash % 7) elements with the ids 1 + index(c) * 16 + slot, inserted in DESCENDING
 * order (insert sorts them); every element lists 1..20 ids of elements from its own container and the ones around it,
 * across the seam on a torus, from the edge container beyond a Cube boundary when there is one. */
template<int DIM, bool TORUS>
class MeshInitializer : public SimpleInitializer<ContainerCell<MeshElement<DIM, TORUS>, CONTAINER_CAPACITY> >
{
public:
    typedef MeshElement<DIM, TORUS> Cargo;
    typedef ContainerCell<Cargo, CONTAINER_CAPACITY> Cell;

    MeshInitializer(const Coord<DIM>& dim, unsigned steps, bool edge, int missingID = 0) :
        SimpleInitializer<Cell>(dim, steps), edge(edge), missingID(missingID)
    {}

    static int count(const Coord<DIM>& c, const Coord<DIM>& dim)
    {
        return (int)(splitmix(1000 + c.toIndex(dim)) % 7);
    }

    static int edgeBase(const Coord<DIM>& dim)
    {
        return 1 + (int)dim.prod() * CONTAINER_CAPACITY;
    }

    static Cell edgeCell(const Coord<DIM>& dim)
    {
        Cell e;
        for (int s = 0; s < 5; ++s) {
            e.insert(edgeBase(dim) + s, Cargo(edgeBase(dim) + s, 3.0 + s, 0));
        }
        return e;
    }

    virtual void grid(GridBase<Cell, DIM> *ret)
    {
        Coord<DIM> dim = this->gridDimensions();
        CoordBox<DIM> whole(Coord<DIM>(), dim);
        ret->setEdge(edge ? edgeCell(dim) : Cell());
        CoordBox<DIM> box = ret->boundingBox();
        for (typename CoordBox<DIM>::Iterator i = box.begin(); i != box.end(); ++i) {
            if (!whole.inBounds(*i)) {
                continue;
            }
            Cell cell;
            int n = count(*i, dim);
            for (int s = n - 1; s >= 0; --s) {
                int id = 1 + (int)i->toIndex(dim) * CONTAINER_CAPACITY + s;
                Cargo e(id, uniform(id), (id % 5 == 0) ? 0.125 * uniform(7777 + id) : 0.0);
                int links = 1 + (int)(splitmix(2 * id) % CONTAINER_MAX_NEIGHBORS);
                for (int j = 0; j < links; ++j) {
                    uint64_t r = splitmix(((uint64_t)id << 8) + j);
                    Coord<DIM> o = *i;
                    for (int a = 0; a < DIM; ++a) {
                        o[a] += (int)((r >> (8 * a)) % 3) - 1;
                    }
                    int target;
                    if (TORUS) {
                        o = Wrap<DIM>::torus(o, dim);
                    }
                    if (!whole.inBounds(o)) {
                        target = edge ? edgeBase(dim) + (int)((r >> 32) % 5) : id;
                    } else if (count(o, dim) == 0) {
                        target = id;
                    } else {
                        target = 1 + (int)o.toIndex(dim) * CONTAINER_CAPACITY + (int)((r >> 32) % count(o, dim));
                    }
                    e.neighborIDs << target;
                }
                if (id == missingID) {
                    e.neighborIDs[0] = 4711;
                }
                cell.insert(id, e);
            }
            ret->set(*i, cell);
        }
    }

private:
    bool edge;
    int missingID;
};

template<int DIM, bool TORUS>
static void compare(const char *name, const Coord<DIM>& dim, unsigned steps, bool edge)
{
    typedef MeshInitializer<DIM, TORUS> Init;
    typedef typename Init::Cell Cell;
    SerialSimulator<Cell> ref(new Init(dim, steps, edge));
    B200Simulator<Cell> dev(new Init(dim, steps, edge));
    ref.run();
    dev.run();
    CHECK(ref.getStep() == steps && dev.getStep() == steps);
    const GridBase<Cell, DIM> *a = ref.getGrid();
    const GridBase<Cell, DIM> *b = dev.getGrid();
    CHECK(a->boundingBox() == b->boundingBox());
    CHECK(a->getEdge().size() == b->getEdge().size());
    std::size_t elements = 0, bad = 0, changed = 0;
    CoordBox<DIM> box(Coord<DIM>(), dim);
    for (typename CoordBox<DIM>::Iterator i = box.begin(); i != box.end(); ++i) {
        Cell ca = a->get(*i), cb = b->get(*i);
        if (ca.size() != cb.size()) {
            ++bad;
            continue;
        }
        elements += ca.size();
        for (std::size_t s = 0; s < ca.size(); ++s) {
            const typename Init::Cargo& ea = ca.begin()[s];
            const typename Init::Cargo& eb = cb.begin()[s];
            if (ca.getIDs()[s] != cb.getIDs()[s] || ea.id != eb.id || std::memcmp(&ea.temperature, &eb.temperature, 8) ||
                ea.influx != eb.influx || !(ea.neighborIDs == eb.neighborIDs)) {
                ++bad;
            }
            changed += ea.temperature != uniform(ea.id);
        }
    }
    CHECK(bad == 0);
    CHECK(elements > 0 && changed > 0);
    std::printf("%-18s %s x %2u steps: %zu elements, %s\n", name, dim.toString().c_str(), steps, elements,
                bad == 0 ? "bit-identical to SerialSimulator" : "DIFFERENT");
}

/* a Steerer that heats every element of one container every third step, and a Writer that sums all temperatures at
 * every step: both plugins see ContainerCells through GridBase, as with the reference's simulator */
template<typename CELL, int DIM>
class HeatSteerer : public Steerer<CELL>
{
public:
    typedef typename Steerer<CELL>::GridType GridType;
    typedef typename Steerer<CELL>::SteererFeedback SteererFeedback;
    typedef typename Steerer<CELL>::CoordType CoordType;

    HeatSteerer(unsigned period, const Coord<DIM>& where) : Steerer<CELL>(period), where(where) {}

    virtual void nextStep(GridType *grid, const Region<DIM>&, const CoordType&, unsigned step, SteererEvent event, std::size_t, bool,
                          SteererFeedback *)
    {
        if (event != STEERER_NEXT_STEP) {
            return;
        }
        CELL cell = grid->get(where);
        for (typename CELL::Iterator e = cell.begin(); e != cell.end(); ++e) {
            e->temperature += 1.0 + step;
        }
        grid->set(where, cell);
    }

private:
    Coord<DIM> where;
};

template<typename CELL, int DIM>
class SumWriter : public Clonable<Writer<CELL>, SumWriter<CELL, DIM> >
{
public:
    typedef typename Writer<CELL>::GridType GridType;

    SumWriter(std::vector<double> *log, unsigned period) : Clonable<Writer<CELL>, SumWriter<CELL, DIM> >("", period), log(log) {}

    virtual void stepFinished(const GridType& grid, unsigned step, WriterEvent)
    {
        double sum = 0;
        CoordBox<DIM> box = grid.boundingBox();
        for (typename CoordBox<DIM>::Iterator i = box.begin(); i != box.end(); ++i) {
            CELL cell = grid.get(*i);
            for (typename CELL::Iterator e = cell.begin(); e != cell.end(); ++e) {
                sum += e->temperature;
            }
        }
        log->push_back(step);
        log->push_back(sum);
    }

private:
    std::vector<double> *log;
};

static void testPlugins()
{
    typedef MeshInitializer<2, true> Init;
    typedef Init::Cell Cell;
    Coord<2> dim(9, 8);
    unsigned steps = 11;
    std::vector<double> logRef, logDev;
    SerialSimulator<Cell> ref(new Init(dim, steps, false));
    B200Simulator<Cell> dev(new Init(dim, steps, false));
    ref.addSteerer(new HeatSteerer<Cell, 2>(3, Coord<2>(4, 4)));
    dev.addSteerer(new HeatSteerer<Cell, 2>(3, Coord<2>(4, 4)));
    ref.addWriter(new SumWriter<Cell, 2>(&logRef, 2));
    dev.addWriter(new SumWriter<Cell, 2>(&logDev, 2));
    ref.run();
    dev.run();
    CHECK(logRef.size() == logDev.size() && logRef.size() >= 12);
    CHECK(!logRef.empty() && std::memcmp(logRef.data(), logDev.data(), logRef.size() * sizeof(double)) == 0);
    CHECK(logRef.back() != logRef[1]);
    std::printf("Steerer writes every third step + Writer every second: %zu log entries, %s\n", logRef.size() / 2,
                (logRef.size() == logDev.size() && !std::memcmp(logRef.data(), logDev.data(), logRef.size() * sizeof(double)))
                    ? "identical sums at every event" : "DIFFERENT");
}

static void testMissingID()
{
    typedef MeshInitializer<2, false> Init;
    typedef Init::Cell Cell;
    Coord<2> dim(6, 5);
    int id = 0;
    CoordBox<2> box(Coord<2>(), dim);
    for (CoordBox<2>::Iterator i = box.begin(); i != box.end() && !id; ++i) {
        if (Init::count(*i, dim) > 0) {
            id = 1 + (int)i->toIndex(dim) * CONTAINER_CAPACITY;
        }
    }
    bool refThrew = false, devThrew = false;
    std::string what;
    try {
        SerialSimulator<Cell> ref(new Init(dim, 2, false, id));
        ref.run();
    } catch (const std::logic_error&) {
        refThrew = true;
    }
    try {
        B200Simulator<Cell> dev(new Init(dim, 2, false, id));
        dev.run();
    } catch (const std::logic_error& e) {
        devThrew = true;
        what = e.what();
    }
    CHECK(refThrew && devThrew);
    CHECK(what.find("4711") != std::string::npos);
    std::printf("id not found: std::logic_error from both (%s)\n", what.c_str());
}

static void testGridAccess()
{
    typedef MeshElement<2, false> Cargo;
    typedef ContainerCell<Cargo, CONTAINER_CAPACITY> Cell;
    Coord<2> dim(4, 3);
    B200ContainerGrid<Cargo, CONTAINER_CAPACITY> grid(CoordBox<2>(Coord<2>(), dim));
    CHECK(grid.boundingBox() == CoordBox<2>(Coord<2>(), dim));
    CHECK(grid.get(Coord<2>(1, 1)).size() == 0);
    Cell a, b;
    Cargo e1(12, 1.0, 0.5), e2(10, 2.0, 0), e3(11, 4.0, 0);
    e1.neighborIDs << 10 << 11 << 20;
    e2.neighborIDs << 12;
    e3.neighborIDs << 20 << 20;
    a.insert(12, e1);
    a.insert(10, e2);
    a.insert(11, e3);
    Cargo f(20, 8.0, 0);
    f.neighborIDs << 10;
    b.insert(20, f);
    grid.set(Coord<2>(1, 1), a);
    grid.set(Coord<2>(2, 2), b);
    Cell back = grid.get(Coord<2>(1, 1));
    CHECK(back.size() == 3 && back.getIDs()[0] == 10 && back.getIDs()[1] == 11 && back.getIDs()[2] == 12);
    CHECK(back[12]->temperature == 1.0 && back[12]->neighborIDs.size() == 3 && back[12]->id == 12);
    grid.update(0, 1);
    back = grid.get(Coord<2>(1, 1));
    CHECK(back[12]->temperature == 0.5 + (2.0 + 4.0 + 8.0) / 3);
    CHECK(back[10]->temperature == 1.0 && back[11]->temperature == 8.0);
    CHECK(grid.get(Coord<2>(2, 2))[20]->temperature == 2.0);
    // a steering write between sweeps: only container (2, 2) is rewritten, (1, 1) keeps its temperatures
    f.temperature = 100.0;
    b.insert(20, f);
    grid.set(Coord<2>(2, 2), b);
    back = grid.get(Coord<2>(1, 1));
    CHECK(back[10]->temperature == 1.0 && back[11]->temperature == 8.0);
    grid.update(1, 1);
    back = grid.get(Coord<2>(1, 1));
    CHECK(back[11]->temperature == 100.0);
    CHECK(back[12]->temperature == 0.5 + (1.0 + 8.0 + 100.0) / 3);
    // an element from two containers away is out of reach
    Cell far;
    Cargo g(30, 1.0, 0);
    g.neighborIDs << 10;
    far.insert(30, g);
    grid.set(Coord<2>(3, 1), far);
    bool threw = false;
    try {
        grid.update(2, 1);
    } catch (const std::logic_error&) {
        threw = true;
    }
    CHECK(threw);
    // beyond a Cube boundary GridBase::get hands out the edge cell and set() writes it (Topology::locate)
    CHECK(grid.get(Coord<2>(4, 0)).size() == 0 && grid.get(Coord<2>(-1, -1)).size() == 0);
    grid.set(Coord<2>(4, 0), far);
    CHECK(grid.getEdge().size() == 1 && grid.get(Coord<2>(0, 3)).size() == 1 && grid.get(Coord<2>(0, 3))[30] != 0);
    std::vector<Cell> streak(6);
    grid.get(Streak<2>(Coord<2>(-1, 1), 5), streak.data());
    CHECK(streak[0].size() == 1 && streak[2].size() == 3 && streak[5].size() == 1);
    std::printf("B200ContainerGrid set / get / update / steering write: %s\n", failures ? "FAILED" : "ok");
}

int main()
{
    if (b200geo_device_count() < 1) {
        std::printf("no CUDA device: %s\n", b200geo_last_error());
        return 77;
    }
    compare<2, false>("container 2-D cube", Coord<2>(12, 9), 15, true);
    compare<2, false>("container 2-D cube", Coord<2>(31, 2), 6, false);
    compare<2, true>("container 2-D torus", Coord<2>(12, 9), 15, false);
    compare<3, false>("container 3-D cube", Coord<3>(7, 6, 5), 9, true);
    compare<3, true>("container 3-D torus", Coord<3>(7, 6, 5), 9, false);
    compare<3, true>("container 3-D torus", Coord<3>(2, 1, 3), 4, false);
    testPlugins();
    testMissingID();
    testGridAccess();
    if (failures) {
        std::printf("%d checks FAILED\n", failures);
        return 1;
    }
    std::printf("all checks passed\n");
    return 0;
}
