/* Drop-in for grids of ContainerCell containers (ID-keyed cargo: meshfree / unstructured models on the regular
 * grid): B200ContainerGrid<CARGO, SIZE> is a GridBase<ContainerCell<CARGO, SIZE, int>, DIM> whose stepped state lives
 * in the device-resident container grid of libb200geo.so (b200geo_containergrid_*), and B200Simulator<ContainerCell<...> >
 * (b200simulator.h) steps it with the container kernels. Compiles against the UNCHANGED reference headers.
 *
 * Replaces Grid<ContainerCell<CARGO, SIZE> > + ContainerCell::update (storage/containercell.h:24-218) with the ID lookup of
 * NeighborhoodAdapter (storage/neighborhoodadapter.h:45-65) for cargo bound with B200GEO_BIND_CARGO: the binding names the
 * members the bound update reads and writes — the model of src/examples/voronoi/main.cpp:41-54,
 *     TEMPERATURE = INFLUX + (sum over NEIGHBOR_IDS, in list order, of hood[id].TEMPERATURE) / NEIGHBOR_IDS.size();
 * the Python twin is models.ContainerModel. Members the update does not touch (coordinates, shapes, areas ...) never
 * change during a run; the grid keeps them in a host-side copy of the containers and get() returns them with the
 * temperatures from the device. */
#ifndef LIBGEODECOMP_B200_B200CONTAINERGRID_H
#define LIBGEODECOMP_B200_B200CONTAINERGRID_H

#include <libgeodecomp/storage/containercell.h>
#include <libgeodecomp/storage/fixedarray.h>
#include <libgeodecomp/storage/gridbase.h>

#include <cstring>
#include <stdexcept>
#include <type_traits>
#include <vector>

#include "../b200geo.h"

namespace LibGeoDecomp {

template<typename CARGO>
struct B200CargoBinding;   /* specialise with B200GEO_BIND_CARGO */

/* B200GEO_BIND_CARGO(Cargo, temperature, influx, neighborIDs): two double members and a FixedArray<int, N> of Cargo */
#define B200GEO_BIND_CARGO(CARGO, TEMPERATURE, INFLUX, NEIGHBOR_IDS)                                           \
    namespace LibGeoDecomp {                                                                                   \
    template<> struct B200CargoBinding<CARGO> {                                                                \
        typedef std::remove_reference<decltype(((CARGO*)0)->NEIGHBOR_IDS)>::type NeighborList;                 \
        static double& value(CARGO& c) { return c.TEMPERATURE; }                                               \
        static double value(const CARGO& c) { return c.TEMPERATURE; }                                          \
        static double influx(const CARGO& c) { return c.INFLUX; }                                              \
        static const NeighborList& neighbors(const CARGO& c) { return c.NEIGHBOR_IDS; }                        \
        static int maxNeighbors() { return (int)NeighborList::capacity(); }                                    \
    };                                                                                                         \
    }

namespace B200Helpers {
inline void check(int rc);   /* defined in b200simulator.h, which includes this header */
}

template<typename CARGO, std::size_t SIZE>
class B200ContainerGrid : public GridBase<ContainerCell<CARGO, SIZE, int>, APITraits::SelectTopology<CARGO>::Value::DIM>
{
public:
    typedef ContainerCell<CARGO, SIZE, int> CELL;
    typedef B200CargoBinding<CARGO> Binding;
    typedef typename APITraits::SelectTopology<CARGO>::Value Topology;
    static const int DIM = Topology::DIM;
    typedef GridBase<CELL, DIM> Base;

    explicit B200ContainerGrid(const CoordBox<DIM>& box = CoordBox<DIM>(), const CELL& edgeCell = CELL(), int device = 0) :
        Base(box.dimensions),
        box(box),
        edgeCell(edgeCell),
        device(device),
        handle(0)
    {
        static_assert(DIM == 2 || DIM == 3, "ContainerCell grids on the device are 2-D or 3-D");
        create();
    }

    virtual ~B200ContainerGrid()
    {
        b200geo_containergrid_destroy(handle);
    }

    b200geo_containergrid *raw()
    {
        flush();
        return handle;
    }

    virtual void resize(const CoordBox<DIM>& newBox)
    {
        b200geo_containergrid_destroy(handle);
        handle = 0;
        box = newBox;
        this->topoDimensions = newBox.dimensions;
        create();
    }

    virtual void set(const Coord<DIM>& coord, const CELL& cell)
    {
        set(Streak<DIM>(coord, coord.x() + 1), &cell);
    }

    /* Writes land in the host-side copy and are shipped as ONE box (the bounding box of everything written since the
     * last flush) before the next sweep or read: an Initializer that sets container after container
     * (src/examples/voronoi/main.cpp:145-171) costs one transfer, not one per container. */
    virtual void set(const Streak<DIM>& streak, const CELL *cells)
    {
        int n = streak.length();
        for (int i = 0; i < n; ++i) {
            Coord<DIM> c = streak.origin;
            c.x() += i;
            // coordinates beyond the box address what Grid::operator[] does (Topology::locate, geometry/topologies.h:186-213):
            // the periodic image on a Torus axis, the edge cell beyond a Cube boundary
            if (!locate(&c)) {
                setEdge(cells[i]);
                continue;
            }
            std::size_t at = index(c);
            shadow[at] = cells[i];
            written[at] = 1;
            touch(c);
        }
    }

    virtual CELL get(const Coord<DIM>& coord) const
    {
        CELL cell;
        get(Streak<DIM>(coord, coord.x() + 1), &cell);
        return cell;
    }

    virtual void get(const Streak<DIM>& streak, CELL *cells) const
    {
        const_cast<B200ContainerGrid*>(this)->flush();
        int n = streak.length();
        if (n <= 0) {
            return;
        }
        if (!box.inBounds(streak)) {
            // cell by cell like Grid::get (storage/grid.h:225-240): periodic images and the edge cell
            for (int i = 0; i < n; ++i) {
                Coord<DIM> c = streak.origin;
                c.x() += i;
                if (locate(&c)) {
                    get(Streak<DIM>(c, c.x() + 1), cells + i);
                } else {
                    cells[i] = edgeCell;
                }
            }
            return;
        }
        std::vector<double> values((std::size_t)n * SIZE);
        b200geo_container_box b;
        std::memset(&b, 0, sizeof(b));
        b.values = values.data();
        int32_t o[3], d[3] = {n, 1, 1};
        local(streak.origin, o);
        B200Helpers::check(b200geo_containergrid_save(handle, o, d, &b, B200GEO_HOST, 0));
        for (int i = 0; i < n; ++i) {
            Coord<DIM> c = streak.origin;
            c.x() += i;
            cells[i] = shadow[index(c)];
            std::size_t s = 0;
            for (typename CELL::Iterator e = cells[i].begin(); e != cells[i].end(); ++e, ++s) {
                Binding::value(*e) = values[(std::size_t)i * SIZE + s];
            }
        }
    }

    virtual void setEdge(const CELL& cell)
    {
        edgeCell = cell;
        Packed p(1, Binding::maxNeighbors());
        p.put(0, cell);
        b200geo_container_box b = p.box();
        B200Helpers::check(b200geo_containergrid_set_edge(handle, &b));
    }

    virtual const CELL& getEdge() const
    {
        return edgeCell;
    }

    virtual CoordBox<DIM> boundingBox() const
    {
        return box;
    }

    /* n sweeps: every cargo of every container updated against the old grid; the first sweep after a write resolves
     * the neighbour IDs (std::logic_error "id not found" as from NeighborhoodAdapter::operator[]) */
    void update(unsigned firstNanoStep, unsigned sweeps)
    {
        flush();
        B200Helpers::check(b200geo_containergrid_step(handle, firstNanoStep, sweeps, 0));
    }

    void sync() const
    {
        B200Helpers::check(b200geo_sync(0));
    }

protected:
    virtual void saveMemberImplementation(char *, MemoryLocation::Location, const Selector<CELL>&,
                                          const typename Region<DIM>::StreakIterator&,
                                          const typename Region<DIM>::StreakIterator&) const
    {
        throw std::logic_error("B200ContainerGrid: containers have no selectable members");
    }

    virtual void loadMemberImplementation(const char *, MemoryLocation::Location, const Selector<CELL>&,
                                          const typename Region<DIM>::StreakIterator&,
                                          const typename Region<DIM>::StreakIterator&)
    {
        throw std::logic_error("B200ContainerGrid: containers have no selectable members");
    }

private:
    /* a box of containers in the interchange format of include/b200geo.h */
    struct Packed {
        std::vector<int32_t> counts, ids, nbCounts, nbIDs;
        std::vector<double> values, influx;
        int maxNB;

        Packed(std::size_t cells, int maxNB) :
            counts(cells, 0), ids(cells * SIZE, 0), nbCounts(cells * SIZE, 0), nbIDs(cells * SIZE * maxNB, 0),
            values(cells * SIZE, 0), influx(cells * SIZE, 0), maxNB(maxNB)
        {}

        void put(std::size_t i, const CELL& cell)
        {
            counts[i] = (int32_t)cell.size();
            std::size_t s = 0;
            for (typename CELL::ConstIterator e = cell.begin(); e != cell.end(); ++e, ++s) {
                std::size_t slot = i * SIZE + s;
                ids[slot] = cell.getIDs()[s];
                values[slot] = Binding::value(*e);
                influx[slot] = Binding::influx(*e);
                const typename Binding::NeighborList& nb = Binding::neighbors(*e);
                nbCounts[slot] = (int32_t)nb.size();
                for (std::size_t j = 0; j < nb.size(); ++j) {
                    nbIDs[slot * maxNB + j] = nb[j];
                }
            }
        }

        b200geo_container_box box()
        {
            b200geo_container_box b;
            b.counts = counts.data();
            b.ids = ids.data();
            b.values = values.data();
            b.influx = influx.data();
            b.nb_counts = nbCounts.data();
            b.nb_ids = nbIDs.data();
            return b;
        }
    };

    CoordBox<DIM> box;
    CELL edgeCell;
    int device;
    b200geo_containergrid *handle;
    std::vector<CELL> shadow;        /* every container as last written; the temperatures of record are the device's */
    bool dirty;
    int32_t dirtyLo[3], dirtyHi[3];  /* bounding box (local coordinates, half open) of the containers written since the last flush */

    /* Topology::locate for a coordinate: false = the edge cell, else *c is moved to the cell inside the box */
    bool locate(Coord<DIM> *c) const
    {
        Coord<DIM> rel = *c - box.origin;
        if (Topology::isOutOfBounds(rel, box.dimensions)) {
            return false;
        }
        *c = Topology::normalize(rel, box.dimensions) + box.origin;
        return true;
    }

    void local(const Coord<DIM>& c, int32_t *o) const
    {
        o[0] = o[1] = o[2] = 0;
        for (int i = 0; i < DIM; ++i) {
            o[i] = c[i] - box.origin[i];
        }
    }

    std::size_t index(const Coord<DIM>& c) const
    {
        int32_t o[3];
        local(c, o);
        for (int i = 0; i < DIM; ++i) {
            if (o[i] < 0 || o[i] >= box.dimensions[i]) {
                throw std::invalid_argument("B200ContainerGrid: coordinate outside the grid");
            }
        }
        std::size_t ny = DIM > 1 ? box.dimensions[1] : 1;
        return ((std::size_t)o[2] * ny + o[1]) * box.dimensions[0] + o[0];
    }

    void touch(const Coord<DIM>& c)
    {
        int32_t o[3];
        local(c, o);
        for (int i = 0; i < 3; ++i) {
            if (!dirty || o[i] < dirtyLo[i]) dirtyLo[i] = o[i];
            if (!dirty || o[i] + 1 > dirtyHi[i]) dirtyHi[i] = o[i] + 1;
        }
        dirty = true;
    }

    void flush()
    {
        if (!dirty) {
            return;
        }
        int32_t d[3];
        std::size_t cells = 1;
        for (int i = 0; i < 3; ++i) {
            d[i] = dirtyHi[i] - dirtyLo[i];
            cells *= d[i];
        }
        // containers of the box that were not written keep their temperatures: fetch them before the box is rewritten
        std::vector<double> current(cells * SIZE);
        b200geo_container_box cur;
        std::memset(&cur, 0, sizeof(cur));
        cur.values = current.data();
        B200Helpers::check(b200geo_containergrid_save(handle, dirtyLo, d, &cur, B200GEO_HOST, 0));
        Packed p(cells, Binding::maxNeighbors());
        std::size_t nx = box.dimensions[0], ny = DIM > 1 ? box.dimensions[1] : 1, i = 0;
        for (int z = dirtyLo[2]; z < dirtyHi[2]; ++z) {
            for (int y = dirtyLo[1]; y < dirtyHi[1]; ++y) {
                for (int x = dirtyLo[0]; x < dirtyHi[0]; ++x, ++i) {
                    std::size_t at = ((std::size_t)z * ny + y) * nx + x;
                    p.put(i, shadow[at]);
                    if (!written[at]) {
                        std::memcpy(&p.values[i * SIZE], &current[i * SIZE], SIZE * sizeof(double));
                    }
                    written[at] = 0;
                }
            }
        }
        b200geo_container_box b = p.box();
        B200Helpers::check(b200geo_containergrid_load(handle, dirtyLo, d, &b, B200GEO_HOST, 0));
        dirty = false;
    }

    std::vector<char> written;       /* containers written since the last flush */

    void create()
    {
        b200geo_containergrid_desc desc;
        std::memset(&desc, 0, sizeof(desc));
        desc.n_dims = DIM;
        std::size_t cells = 1;
        for (int i = 0; i < 3; ++i) {
            desc.dim[i] = i < DIM ? box.dimensions[i] : 1;
            cells *= desc.dim[i];
            int mode = (i < DIM && Topology::wrapsAxis(i)) ? B200GEO_GHOST_WRAP : B200GEO_GHOST_EDGE;
            desc.ghost_mode[i][0] = desc.ghost_mode[i][1] = mode;
        }
        desc.capacity = (int32_t)SIZE;
        desc.max_neighbors = Binding::maxNeighbors();
        B200Helpers::check(b200geo_containergrid_create(&desc, device, &handle));
        shadow.assign(cells, CELL());
        written.assign(cells, 0);
        dirty = false;
        setEdge(edgeCell);
    }
};

}

#endif
