/* Drop-in for grids of BoxCell containers (short-range n-body): B200BoxGrid<PARTICLE, N> is a
 * GridBase<BoxCell<FixedArray<PARTICLE, N> >, 3> whose storage is the device-resident container
 * grid of libb200geo.so (b200geo_boxgrid_*), and B200Simulator<BoxCell<...> > (b200simulator.h)
 * steps it with the n-body kernels. Compiles against the UNCHANGED reference headers.
 *
 * Replaces Grid<BoxCell<FixedArray<PARTICLE, N> > > + BoxCell::update (storage/boxcell.h:21-178) for
 * particles bound with B200GEO_BIND_PARTICLE: the binding says where position and velocity live
 * inside the particle and supplies the model constants, the Python twin is models.NBodyModel.
 * A container's origin and dimension follow from its coordinate (origin = coord * cellEdge), which
 * is how the bound model's Initializer must construct them; set() checks the dimension. */
#ifndef LIBGEODECOMP_B200_B200BOXGRID_H
#define LIBGEODECOMP_B200_B200BOXGRID_H

#include <libgeodecomp/storage/boxcell.h>
#include <libgeodecomp/storage/fixedarray.h>
#include <libgeodecomp/storage/gridbase.h>

#include <cstring>
#include <stdexcept>
#include <vector>

#include "../b200geo.h"

namespace LibGeoDecomp {

template<typename PARTICLE>
struct B200ParticleBinding;   /* specialise with B200GEO_BIND_PARTICLE */

/* B200GEO_BIND_PARTICLE(Particle, float, pos, vel, dtExpr, cutoffExpr, cellEdgeExpr): pos and vel are
 * REAL[3] members of Particle */
#define B200GEO_BIND_PARTICLE(PARTICLE, REAL_TYPE, POS, VEL, DT, CUTOFF, CELL_EDGE)      \
    namespace LibGeoDecomp {                                                            \
    template<> struct B200ParticleBinding<PARTICLE> {                                   \
        typedef REAL_TYPE Real;                                                         \
        static const Real *pos(const PARTICLE& p) { return p.POS; }                     \
        static const Real *vel(const PARTICLE& p) { return p.VEL; }                     \
        static Real *pos(PARTICLE& p) { return p.POS; }                                 \
        static Real *vel(PARTICLE& p) { return p.VEL; }                                 \
        static double dt() { return DT; }                                               \
        static double cutoff() { return CUTOFF; }                                       \
        static double cellEdge() { return CELL_EDGE; }                                  \
    };                                                                                  \
    }

namespace B200Helpers {
inline void check(int rc);   /* defined in b200simulator.h, which includes this header */
}

template<typename PARTICLE, int N>
class B200BoxGrid : public GridBase<BoxCell<FixedArray<PARTICLE, N> >, 3>
{
public:
    typedef BoxCell<FixedArray<PARTICLE, N> > CELL;
    typedef B200ParticleBinding<PARTICLE> Binding;
    typedef typename Binding::Real Real;
    static const int DIM = 3;
    typedef GridBase<CELL, 3> Base;

    explicit B200BoxGrid(const CoordBox<3>& box = CoordBox<3>(), const CELL& edgeCell = CELL(), int device = 0) :
        Base(box.dimensions),
        box(box),
        edgeCell(edgeCell),
        device(device),
        handle(0),
        lowPeer(false),
        highPeer(false)
    {
        create();
    }

    /* one slab (along z) of a larger container grid: a face towards a neighbouring slab is a ghost plane of
     * that neighbour's containers (B200StripingSimulator) */
    B200BoxGrid(const CoordBox<3>& box, const CELL& edgeCell, int device, bool lowPeer, bool highPeer) :
        Base(box.dimensions),
        box(box),
        edgeCell(edgeCell),
        device(device),
        handle(0),
        lowPeer(lowPeer),
        highPeer(highPeer)
    {
        create();
    }

    b200geo_boxgrid *raw()
    {
        return handle;
    }

    static b200geo_nbody_params parameters()
    {
        b200geo_nbody_params p;
        p.dt = Binding::dt();
        p.cutoff = Binding::cutoff();
        p.nano_steps = APITraits::SelectNanoSteps<PARTICLE>::VALUE;
        return p;
    }

    virtual ~B200BoxGrid()
    {
        b200geo_boxgrid_destroy(handle);
    }

    virtual void resize(const CoordBox<3>& newBox)
    {
        b200geo_boxgrid_destroy(handle);
        handle = 0;
        box = newBox;
        this->topoDimensions = newBox.dimensions;
        create();
    }

    virtual void set(const Coord<3>& coord, const CELL& cell)
    {
        set(Streak<3>(coord, coord.x() + 1), &cell);
    }

    virtual void set(const Streak<3>& streak, const CELL *cells)
    {
        int n = streak.length();
        std::vector<int32_t> counts(n);
        std::vector<Real> parts((std::size_t)n * N * 6, 0);
        for (int i = 0; i < n; ++i) {
            const CELL& c = cells[i];
            counts[i] = (int32_t)c.size();
            for (std::size_t p = 0; p < c.size(); ++p) {
                Real *dst = &parts[((std::size_t)i * N + p) * 6];
                for (int k = 0; k < 3; ++k) {
                    dst[k] = Binding::pos(c[p])[k];
                    dst[3 + k] = Binding::vel(c[p])[k];
                }
            }
        }
        int32_t o[3], d[3] = {n, 1, 1};
        local(streak.origin, o);
        B200Helpers::check(b200geo_boxgrid_load(handle, o, d, counts.data(), parts.data(), B200GEO_HOST, 1, 0));
    }

    virtual CELL get(const Coord<3>& coord) const
    {
        CELL cell;
        get(Streak<3>(coord, coord.x() + 1), &cell);
        return cell;
    }

    virtual void get(const Streak<3>& streak, CELL *cells) const
    {
        int n = streak.length();
        std::vector<int32_t> counts(n);
        std::vector<Real> parts((std::size_t)n * N * 6);
        int32_t o[3], d[3] = {n, 1, 1};
        local(streak.origin, o);
        // a container that overflowed on the device is the reference's std::out_of_range (fixedarray.h:77-83)
        B200Helpers::check(b200geo_boxgrid_check(handle, 0));
        B200Helpers::check(b200geo_boxgrid_save(handle, o, d, counts.data(), parts.data(), B200GEO_HOST, 0));
        double e = Binding::cellEdge();
        for (int i = 0; i < n; ++i) {
            Coord<3> c = streak.origin + Coord<3>(i, 0, 0);
            cells[i] = CELL(FloatCoord<3>(c.x() * e, c.y() * e, c.z() * e), FloatCoord<3>(e, e, e));
            for (int p = 0; p < counts[i]; ++p) {
                PARTICLE particle;
                const Real *src = &parts[((std::size_t)i * N + p) * 6];
                for (int k = 0; k < 3; ++k) {
                    Binding::pos(particle)[k] = src[k];
                    Binding::vel(particle)[k] = src[3 + k];
                }
                cells[i] << particle;
            }
        }
    }

    virtual void setEdge(const CELL& cell)
    {
        if (cell.size() != 0) {
            throw std::logic_error("B200BoxGrid: the edge container of a Cube grid must be empty");
        }
        edgeCell = cell;
    }

    virtual const CELL& getEdge() const
    {
        return edgeCell;
    }

    virtual CoordBox<3> boundingBox() const
    {
        return box;
    }

    /* n sweeps: re-bin (nanoStep % NANO_STEPS == 0) or copy, update every particle; swap */
    void update(unsigned firstNanoStep, unsigned sweeps)
    {
        b200geo_nbody_params p = parameters();
        B200Helpers::check(b200geo_boxgrid_step(handle, &p, firstNanoStep, sweeps, 0));
    }

    void sync() const
    {
        B200Helpers::check(b200geo_sync(0));
    }

protected:
    virtual void saveMemberImplementation(char *, MemoryLocation::Location, const Selector<CELL>&,
                                          const typename Region<3>::StreakIterator&,
                                          const typename Region<3>::StreakIterator&) const
    {
        throw std::logic_error("B200BoxGrid: containers have no selectable members");
    }

    virtual void loadMemberImplementation(const char *, MemoryLocation::Location, const Selector<CELL>&,
                                          const typename Region<3>::StreakIterator&,
                                          const typename Region<3>::StreakIterator&)
    {
        throw std::logic_error("B200BoxGrid: containers have no selectable members");
    }

private:
    CoordBox<3> box;
    CELL edgeCell;
    int device;
    b200geo_boxgrid *handle;
    bool lowPeer;
    bool highPeer;

    void local(const Coord<3>& c, int32_t *o) const
    {
        for (int i = 0; i < 3; ++i) {
            o[i] = c[i] - box.origin[i];
        }
    }

    void create()
    {
        b200geo_boxgrid_desc desc;
        std::memset(&desc, 0, sizeof(desc));
        for (int i = 0; i < 3; ++i) {
            desc.dim[i] = box.dimensions[i];
            desc.ghost_mode[i][0] = desc.ghost_mode[i][1] = B200GEO_GHOST_EDGE;
            desc.cell_origin[i] = box.origin[i];
        }
        if (lowPeer) {
            desc.ghost_mode[2][0] = B200GEO_GHOST_PEER;
        }
        if (highPeer) {
            desc.ghost_mode[2][1] = B200GEO_GHOST_PEER;
        }
        desc.capacity = N;
        desc.real_bytes = (int)sizeof(Real);
        desc.cell_edge = Binding::cellEdge();
        B200Helpers::check(b200geo_boxgrid_create(&desc, device, &handle));
    }
};

}

#endif
