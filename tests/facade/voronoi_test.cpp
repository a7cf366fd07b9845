/* The reference's own Voronoi example (src/examples/voronoi/main.cpp: SimpleCell in ContainerCell<SimpleCell, 1000>,
 * VoronoiInitializer over geometry/voronoimesher.h), compiled UNCHANGED from where it lies, run through the reference's
 * SerialSimulator and through B200Simulator: the user adds one B200GEO_BIND_CARGO line and swaps the simulator type.
 * Every element of every container is compared after the run: temperature bit for bit, and the members the bound
 * update never touches (centre, area, shape, neighbour lists, border lengths) unchanged.
 * The example's main() registers a SiloWriter (needs the Silo library, not in this image); main() is compiled but never
 * called, with a stand-in class of that name so that it compiles. */
#include <libgeodecomp.h>

namespace LibGeoDecomp {

template<typename CELL>
class SiloWriter : public Clonable<Writer<CELL>, SiloWriter<CELL> >
{
public:
    SiloWriter(const std::string& prefix, unsigned period) : Clonable<Writer<CELL>, SiloWriter<CELL> >(prefix, period) {}
    template<typename MEMBER, typename CARGO>
    void addSelectorForPointMesh(MEMBER CARGO::*, const std::string&) {}
    virtual void stepFinished(const typename Writer<CELL>::GridType&, unsigned, WriterEvent) {}
};

}

#define main voronoi_example_main
#include <examples/voronoi/main.cpp>
#undef main

#include <libgeodecomp_b200/b200simulator.h>

/* the one line a user adds: which members the element's update() reads and writes */
B200GEO_BIND_CARGO(SimpleCell<FloatCoord>, temperature, influx, neighborIDs)

#include <cstdio>
#include <cstring>

static int failures = 0;
#define CHECK(COND)                                                                     \
    do {                                                                                \
        if (!(COND)) {                                                                  \
            ++failures;                                                                 \
            std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #COND);              \
        }                                                                               \
    } while (0)

template<typename SIM>
static void runExample(SIM **sim, const Coord<2>& dim, unsigned steps, std::size_t numCells)
{
    // the mesher draws the element positions from the library's global generator, seeded per container
    // (geometry/voronoimesher.h:61-77): both simulators see the same mesh
    *sim = new SIM(new VoronoiInitializer(dim, steps, numCells, 400, 100));
    (*sim)->run();
}

static void compare(const Coord<2>& dim, unsigned steps, std::size_t numCells)
{
    SerialSimulator<ContainerCellType> *ref = 0;
    B200Simulator<ContainerCellType> *dev = 0;
    runExample(&ref, dim, steps, numCells);
    runExample(&dev, dim, steps, numCells);
    CHECK(ref->getStep() == steps && dev->getStep() == steps);
    const GridBase<ContainerCellType, 2> *a = ref->getGrid();
    const GridBase<ContainerCellType, 2> *b = dev->getGrid();
    std::size_t elements = 0, bad = 0, warm = 0, links = 0;
    CoordBox<2> box(Coord<2>(), dim);
    for (CoordBox<2>::Iterator i = box.begin(); i != box.end(); ++i) {
        ContainerCellType ca = a->get(*i), cb = b->get(*i);
        if (ca.size() != cb.size()) {
            ++bad;
            continue;
        }
        elements += ca.size();
        for (std::size_t s = 0; s < ca.size(); ++s) {
            const SimpleCell<FloatCoord>& ea = ca.begin()[s];
            const SimpleCell<FloatCoord>& eb = cb.begin()[s];
            bool same = ca.getIDs()[s] == cb.getIDs()[s] && ea.id == eb.id &&
                !std::memcmp(&ea.temperature, &eb.temperature, sizeof(double)) && ea.influx == eb.influx &&
                ea.center == eb.center && ea.area == eb.area && ea.shape == eb.shape && ea.neighborIDs == eb.neighborIDs &&
                ea.neighborBorderLengths == eb.neighborBorderLengths;
            bad += !same;
            warm += ea.temperature > 0;
            links += ea.neighborIDs.size();
        }
    }
    CHECK(bad == 0);
    // (the mesher records an element's OWN id for every Voronoi neighbour, geometry/voronoimesher.h:103-105, so in the
    // reference only the heater of container (0, 0) warms up: one degree per step)
    CHECK(elements > 0 && warm >= 1 && links > elements);
    std::printf("voronoi example %s x %u steps: %zu elements, %zu links, %zu elements warmed up, %s\n", dim.toString().c_str(), steps,
                elements, links, warm, bad == 0 ? "bit-identical to SerialSimulator" : "DIFFERENT");
    delete ref;
    delete dev;
}

int main()
{
    if (b200geo_device_count() < 1) {
        std::printf("no CUDA device: %s\n", b200geo_last_error());
        return 77;
    }
    compare(Coord<2>(10, 5), 60, 100);     // the example's own grid and element count
    compare(Coord<2>(4, 7), 25, 40);
    if (failures) {
        std::printf("%d checks FAILED\n", failures);
        return 1;
    }
    std::printf("all checks passed\n");
    return 0;
}
