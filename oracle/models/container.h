/* TEST INFRASTRUCTURE — an ID-keyed (meshfree / unstructured) model written against the UNCHANGED LibGeoDecomp
 * plugin API: mesh elements live in ContainerCell<MeshElement, SIZE> containers (storage/containercell.h:24-218), one
 * container per cell of the regular grid; ContainerCell::update copies the old container over and calls
 * MeshElement::update(hood, nanoStep) for each element (containercell.h:170-200); hood[id] is
 * NeighborhoodAdapter::operator[] (storage/neighborhoodadapter.h:45-65): the element with that ID from the container
 * itself or, failing that, from the first of the other 3^DIM - 1 containers around it (CoordBox order) that holds it;
 * std::logic_error("id not found") otherwise.
 *
 * The element follows the cell of src/examples/voronoi/main.cpp:9-118 (update: lines 41-54): the new temperature is
 * the influx plus the mean of the neighbours' temperatures, summed in the order of the neighbour list.
 * The same source is the model the CUDA kernel (libgeodecomp_b200/csrc/container.cu) restates.
 *
 * Raw file format of the driver below (whole grid; cells in [nz][ny][nx] order, x fastest), with cap =
 * CONTAINER_CAPACITY and maxnb = CONTAINER_MAX_NEIGHBORS:
 *   int32  counts[cells]   int32 ids[cells][cap]   double values[cells][cap]   double influx[cells][cap]
 *   int32  nb_counts[cells][cap]   int32 nb_ids[cells][cap][maxnb]
 * followed, with --edge, by the same six arrays for ONE container: the edge cell. Output: the six arrays of the grid.
 */
#ifndef B200GEO_ORACLE_MODELS_CONTAINER_H
#define B200GEO_ORACLE_MODELS_CONTAINER_H

#include <libgeodecomp/misc/apitraits.h>
#include <libgeodecomp/storage/containercell.h>
#include <libgeodecomp/storage/fixedarray.h>

namespace b200models {

using namespace LibGeoDecomp;

const int CONTAINER_CAPACITY = 16;
const int CONTAINER_MAX_NEIGHBORS = 20;

template<int DIM, bool TORUS> struct MeshTopology;
template<int DIM> struct MeshTopology<DIM, false> { typedef APITraits::HasCubeTopology<DIM> API; };
template<int DIM> struct MeshTopology<DIM, true> { typedef APITraits::HasTorusTopology<DIM> API; };

template<int DIM, bool TORUS>
class MeshElement
{
public:
    class API : public MeshTopology<DIM, TORUS>::API
    {};

    explicit MeshElement(int id = 0, double temperature = 0, double influx = 0) :
        id(id),
        temperature(temperature),
        influx(influx)
    {}

    template<typename NEIGHBORHOOD>
    void update(const NEIGHBORHOOD& hood, int /* nanoStep */)
    {
        temperature = 0;
        for (FixedArray<int, CONTAINER_MAX_NEIGHBORS>::iterator i = neighborIDs.begin(); i != neighborIDs.end(); ++i) {
            temperature += hood[*i].temperature;
        }
        temperature = influx + temperature / neighborIDs.size();
    }

    int id;
    double temperature;
    double influx;
    FixedArray<int, CONTAINER_MAX_NEIGHBORS> neighborIDs;
};

}

#ifdef B200GEO_ORACLE_REF_DRIVER_H

namespace b200models {

using namespace refdriver;

struct ContainerArrays {
    const int *counts;
    const int *ids;
    const double *values;
    const double *influx;
    const int *nbCounts;
    const int *nbIDs;

    static std::size_t bytes(std::size_t cells)
    {
        const std::size_t cap = CONTAINER_CAPACITY, nb = CONTAINER_MAX_NEIGHBORS;
        return cells * (4 + cap * (4 + 8 + 8 + 4 + 4 * nb));
    }

    ContainerArrays(const char *raw, std::size_t cells)
    {
        const std::size_t cap = CONTAINER_CAPACITY;
        counts = (const int*)raw;
        ids = counts + cells;
        values = (const double*)(ids + cells * cap);
        influx = values + cells * cap;
        nbCounts = (const int*)(influx + cells * cap);
        nbIDs = nbCounts + cells * cap;
    }

    template<typename CONTAINER>
    CONTAINER container(std::size_t idx) const
    {
        typedef typename CONTAINER::Cargo Cargo;
        CONTAINER c;
        for (int s = 0; s < counts[idx]; ++s) {
            std::size_t slot = idx * CONTAINER_CAPACITY + s;
            Cargo e(ids[slot], values[slot], influx[slot]);
            for (int j = 0; j < nbCounts[slot]; ++j) {
                e.neighborIDs << nbIDs[slot * CONTAINER_MAX_NEIGHBORS + j];
            }
            c.insert(ids[slot], e);
        }
        return c;
    }
};

template<int DIM, bool TORUS>
class ContainerInitializer : public SimpleInitializer<ContainerCell<MeshElement<DIM, TORUS>, CONTAINER_CAPACITY> >
{
public:
    typedef ContainerCell<MeshElement<DIM, TORUS>, CONTAINER_CAPACITY> Cell;

    ContainerInitializer(const Coord<DIM>& dim, unsigned steps, const std::vector<char> *raw, bool haveEdge) :
        SimpleInitializer<Cell>(dim, steps), raw(raw), haveEdge(haveEdge)
    {}

    virtual void grid(GridBase<Cell, DIM> *ret)
    {
        Coord<DIM> dim = this->gridDimensions();
        CoordBox<DIM> whole(Coord<DIM>(), dim);
        std::size_t cells = (std::size_t)dim.prod();
        ContainerArrays a(raw->data(), cells);
        if (haveEdge) {
            ContainerArrays e(raw->data() + ContainerArrays::bytes(cells), 1);
            ret->setEdge(e.template container<Cell>(0));
        } else {
            ret->setEdge(Cell());
        }
        CoordBox<DIM> box = ret->boundingBox();
        for (typename CoordBox<DIM>::Iterator i = box.begin(); i != box.end(); ++i) {
            if (whole.inBounds(*i)) {
                ret->set(*i, a.template container<Cell>(i->toIndex(dim)));
            }
        }
    }

private:
    const std::vector<char> *raw;
    bool haveEdge;
};

template<int DIM, bool TORUS, template<typename> class SIM>
int containerRun(int nx, int ny, int nz, unsigned steps, const char *in, const char *out, bool haveEdge, const char *simName, int threads)
{
    typedef ContainerCell<MeshElement<DIM, TORUS>, CONTAINER_CAPACITY> Cell;
    Coord<DIM> dim = Dims<DIM>::make(nx, ny, nz);
    std::size_t cells = (std::size_t)dim.prod();
    std::vector<char> raw = readFile(in);
    std::size_t want = ContainerArrays::bytes(cells) + (haveEdge ? ContainerArrays::bytes(1) : 0);
    if (raw.size() != want) {
        fprintf(stderr, "input size %zu != %zu\n", raw.size(), want);
        return 2;
    }
    SIM<Cell> sim(new ContainerInitializer<DIM, TORUS>(dim, steps, &raw, haveEdge));
    auto t0 = std::chrono::steady_clock::now();
    sim.run();
    auto t1 = std::chrono::steady_clock::now();
    double wall = std::chrono::duration<double>(t1 - t0).count();
    double compute = sim.gatherStatistics()[0].template interval<TimeCompute>();

    const GridBase<Cell, DIM> *grid = sim.getGrid();
    const std::size_t cap = CONTAINER_CAPACITY, nb = CONTAINER_MAX_NEIGHBORS;
    std::vector<char> res(ContainerArrays::bytes(cells), 0);
    ContainerArrays o(res.data(), cells);
    std::size_t elements = 0, links = 0;
    CoordBox<DIM> whole(Coord<DIM>(), dim);
    for (typename CoordBox<DIM>::Iterator i = whole.begin(); i != whole.end(); ++i) {
        Cell cell = grid->get(*i);
        std::size_t idx = i->toIndex(dim);
        const_cast<int*>(o.counts)[idx] = (int)cell.size();
        elements += cell.size();
        for (std::size_t s = 0; s < cell.size(); ++s) {
            const MeshElement<DIM, TORUS>& e = cell.begin()[s];
            std::size_t slot = idx * cap + s;
            const_cast<int*>(o.ids)[slot] = cell.getIDs()[s];
            const_cast<double*>(o.values)[slot] = e.temperature;
            const_cast<double*>(o.influx)[slot] = e.influx;
            const_cast<int*>(o.nbCounts)[slot] = (int)e.neighborIDs.size();
            links += e.neighborIDs.size();
            for (std::size_t j = 0; j < e.neighborIDs.size(); ++j) {
                const_cast<int*>(o.nbIDs)[slot * nb + j] = e.neighborIDs[j];
            }
        }
    }
    if (strcmp(out, "-") != 0) {
        FILE *f = fopen(out, "wb");
        if (!f || fwrite(res.data(), 1, res.size(), f) != res.size()) {
            fprintf(stderr, "cannot write %s\n", out);
            return 3;
        }
        fclose(f);
    }
    printf("{\"model\": \"container\", \"simulator\": \"%s\", \"threads\": %d, \"dims\": [%d, %d, %d], \"n_dims\": %d, \"torus\": %d, "
           "\"steps\": %u, \"elements\": %zu, \"links\": %zu, \"time_compute_s\": %.6f, \"wall_run_s\": %.6f, \"geups_compute\": %.6f}\n",
           simName, threads, nx, ny, nz, DIM, (int)TORUS, steps, elements, links, compute, wall, compute > 0 ? 1e-9 * steps * elements / compute : 0.0);
    return 0;
}

template<template<typename> class SIM>
int containerDispatch(int ndims, bool torus, int nx, int ny, int nz, unsigned steps, const char *in, const char *out, bool edge,
                      const char *simName, int threads)
{
    if (ndims == 2) {
        return torus ? containerRun<2, true, SIM>(nx, ny, nz, steps, in, out, edge, simName, threads)
                     : containerRun<2, false, SIM>(nx, ny, nz, steps, in, out, edge, simName, threads);
    }
    return torus ? containerRun<3, true, SIM>(nx, ny, nz, steps, in, out, edge, simName, threads)
                 : containerRun<3, false, SIM>(nx, ny, nz, steps, in, out, edge, simName, threads);
}

/* usage: lgd_ref_container container nx ny nz steps in.raw out.raw [--dims 2|3] [--torus] [--edge] [--omp] */
inline int containerMain(int argc, char **argv)
{
    if (argc < 8) {
        fprintf(stderr, "usage: %s container nx ny nz steps in.raw out.raw [--dims 2|3] [--torus] [--edge] [--omp]\n", argv[0]);
        return 1;
    }
    int nx = atoi(argv[2]), ny = atoi(argv[3]), nz = atoi(argv[4]);
    unsigned steps = (unsigned)atoi(argv[5]);
    bool omp = false, torus = false, edge = false;
    int ndims = 3;
    for (int i = 8; i < argc; ++i) {
        if (!strcmp(argv[i], "--omp")) omp = true;
        if (!strcmp(argv[i], "--torus")) torus = true;
        if (!strcmp(argv[i], "--edge")) edge = true;
        if (!strcmp(argv[i], "--dims") && i + 1 < argc) ndims = atoi(argv[++i]);
    }
    if ((ndims != 2 && ndims != 3) || (ndims == 2 && nz != 1)) {
        fprintf(stderr, "--dims 2 needs nz = 1\n");
        return 1;
    }
    try {
#ifdef _OPENMP
        if (omp) {
            return containerDispatch<OpenMPSimulator>(ndims, torus, nx, ny, nz, steps, argv[6], argv[7], edge, "OpenMPSimulator",
                                                      omp_get_max_threads());
        }
#endif
        return containerDispatch<SerialSimulator>(ndims, torus, nx, ny, nz, steps, argv[6], argv[7], edge, "SerialSimulator", 1);
    } catch (const std::exception& e) {
        fprintf(stderr, "error: %s\n", e.what());
        return 4;
    }
}

}

#endif

#endif
