/* Header-only C++ façade: the drop-in boundary for existing LibGeoDecomp user code.
 *
 *   B200KernelBinding<CELL>  states which hand-written kernel family implements CELL::update /
 *                            CELL::updateLineX and where CELL's members live (the SoA member table
 *                            LIBFLATARRAY_REGISTER_SOA would generate, in registration order).
 *                            Unbound cells in an nvcc translation unit fall back to the generic
 *                            device path of b200generic.h (their own update() as a kernel).
 *   B200Grid<CELL>           a GridBase<CELL, DIM> (storage/gridbase.h:71-309) backed by the
 *                            device-resident SoA grid of libb200geo.so — what Initializers, Writers
 *                            and Steerers are handed; plays the role of CUDASoAGrid
 *                            (storage/cudasoagrid.h:116).
 *   B200Simulator<CELL>      a MonolithicSimulator<CELL> (parallelization/monolithicsimulator.h:17)
 *                            with the event protocol of SerialSimulator
 *                            (parallelization/serialsimulator.h:48-187); replaces CUDASimulator
 *                            (parallelization/cudasimulator.h:298) for bound cells.
 *
 * Compiles against the UNCHANGED reference headers (-I<reference>/src -I<reference>/lib/libflatarray/include)
 * and include/b200geo.h; links libb200geo.so. User models, Initializers, Writers and Steerers are
 * used as they are; the only addition a user makes is one B200GEO_BIND_CELL line per model.
 * Status codes of the C ABI are mapped to the reference's exceptions.
 */
#ifndef LIBGEODECOMP_B200_B200SIMULATOR_H
#define LIBGEODECOMP_B200_B200SIMULATOR_H

#include <libgeodecomp/io/initializer.h>
#include <libgeodecomp/io/steerer.h>
#include <libgeodecomp/io/writer.h>
#include <libgeodecomp/misc/apitraits.h>
#include <libgeodecomp/parallelization/monolithicsimulator.h>
#include <libgeodecomp/storage/gridbase.h>

#include <cstddef>
#include <algorithm>
#include <cstring>
#include <map>
#include <new>
#include <stdexcept>
#include <string>
#include <vector>

#include "../b200geo.h"

namespace LibGeoDecomp {

/* One entry per SoA member, in registration order. */
struct B200Member {
    std::size_t offsetInCell;  /* offsetof(CELL, member) */
    int bytes;                 /* sizeof(member) */
};

template<typename CELL>
struct B200KernelBinding;      /* specialise with B200GEO_BIND_CELL */

#define B200GEO_MEMBER_ENTRY(CELL, MEMBER) \
    { offsetof(CELL, MEMBER), (int)sizeof(((CELL *)0)->MEMBER) },

/* B200GEO_BIND_CELL(Cell, B200GEO_KERNEL_JACOBI7, B200GEO_MEMBER_ENTRY(Cell, temp)) */
#define B200GEO_BIND_CELL(CELL, KERNEL_ID, ...)                                         \
    namespace LibGeoDecomp {                                                            \
    template<> struct B200KernelBinding<CELL> {                                         \
        static int kernel() { return KERNEL_ID; }                                       \
        static void step(b200geo_grid *g, const int32_t *, unsigned first, unsigned n)  \
        {                                                                               \
            B200Helpers::check(b200geo_step(g, KERNEL_ID, 0, first, n, 0));             \
        }                                                                               \
        static void groupStep(b200geo_group *group, unsigned first, unsigned n)         \
        {                                                                               \
            B200Helpers::check(b200geo_group_step(group, KERNEL_ID, 0, first, n));      \
        }                                                                               \
        /* one sweep over a box: current buffer -> scratch buffer, no swap */           \
        static void updateBox(b200geo_grid *g, unsigned nanoStep, const int32_t origin[3], const int32_t dim[3]) \
        {                                                                               \
            B200Helpers::check(b200geo_update_box(g, KERNEL_ID, 0, nanoStep, origin, dim, 0)); \
        }                                                                               \
        /* n fused sweeps over a box on a stream of the caller (b200geo_update_box_n); `final`: the grid is  \
         * observed afterwards (LBM: density / velocity are stored by the last sweep before an observation) */ \
        static void updateBoxN(b200geo_grid *g, unsigned nanoStep, const int32_t origin[3], const int32_t dim[3], \
                               unsigned sweeps, bool final, void *stream)                \
        {                                                                               \
            int32_t lbmMode = final ? 0 : 2;                                            \
            B200Helpers::check(b200geo_update_box_n(g, KERNEL_ID, (KERNEL_ID) == B200GEO_KERNEL_LBM_D3Q19 ? &lbmMode : 0, \
                                                    nanoStep, origin, dim, sweeps, stream)); \
        }                                                                               \
        static std::vector<B200Member> members()                                        \
        {                                                                               \
            B200Member tab[] = { __VA_ARGS__ };                                         \
            return std::vector<B200Member>(tab, tab + sizeof(tab) / sizeof(tab[0]));    \
        }                                                                               \
    };                                                                                  \
    }

namespace B200Helpers {

inline void check(int rc)
{
    if (rc >= 0) {
        return;
    }
    std::string msg = b200geo_last_error();
    switch (rc) {
    case B200GEO_ERR_INVALID:
        throw std::invalid_argument(msg);
    case B200GEO_ERR_LOGIC:
        throw std::logic_error(msg);
    case B200GEO_ERR_OUT_OF_RANGE:
        throw std::out_of_range(msg);
    case B200GEO_ERR_NOMEM:
        throw std::bad_alloc();
    default:
        throw std::runtime_error(msg.find("CUDA error") == 0 ? msg : "CUDA error: " + msg);
    }
}

/* sweeps a bound kernel family takes per launch (csrc/jacobi_tb.cu, csrc/lbm_tb.cu) and the depth a streamed run
 * uses per level (what is fastest on the device: two for the 27-point and the LBM kernels, four for 6 / 7 points) */
inline unsigned fusedSweeps(int kernel)
{
    switch (kernel) {
    case B200GEO_KERNEL_JACOBI6:
    case B200GEO_KERNEL_JACOBI7:
        return 4;
    case B200GEO_KERNEL_JACOBI27:
    case B200GEO_KERNEL_LBM_D3Q19:
        return 2;
    default:
        return 1;
    }
}

/* members a kernel family never rewrites: every sweep reads them from whichever buffer is current, so an upload has
 * to put them into BOTH buffers (LBM: csrc/lbm.cu leaves `state` and the wall cells' density / velocity alone) */
inline bool invariantMember(int kernel, int member)
{
    return kernel == B200GEO_KERNEL_LBM_D3Q19 && member >= 19;
}

/* a binding may create the device grid itself (the generic SoA path asks for the uniform element layout,
 * b200genericsoa.h); bound cells and word-sliced generic cells take the default layout */
template<typename BINDING>
inline auto createGrid(const b200geo_grid_desc *desc, int device, b200geo_grid **out, int) ->
    decltype(BINDING::createGrid(desc, device, out))
{
    return BINDING::createGrid(desc, device, out);
}

template<typename BINDING>
inline int createGrid(const b200geo_grid_desc *desc, int device, b200geo_grid **out, long)
{
    return b200geo_grid_create(desc, device, out);
}

template<int DIM>
inline void toStreak4(const Streak<DIM>& s, const Coord<DIM>& origin, int32_t *out)
{
    out[0] = s.origin[0] - origin[0];
    out[1] = DIM > 1 ? s.origin[1] - origin[1] : 0;
    out[2] = DIM > 2 ? s.origin[2] - origin[2] : 0;
    out[3] = s.endX - origin[0];
}


/* A run of streaks that tile a box: rows of equal extent in x at consecutive y, planes of equal extent in
 * (x, y) at consecutive z. A dense [z][y][x] array of the box IS the concatenation of its streaks, so Selector
 * I/O and region I/O over such a run is ONE strided copy (b200geo_grid_save_member / _load_member) instead of
 * one call per streak. Box-shaped regions — a Writer's whole grid, a slab, a ghost zone face — collapse to a
 * single box (geometry/region.h:582 keeps them as run-length streaks; SURVEY App. B asks for this fast path). */
struct StreakBox {
    int32_t origin[3];   /* relative to the grid's origin */
    int32_t dim[3];
    std::size_t cells() const { return (std::size_t)dim[0] * dim[1] * dim[2]; }
};

template<int DIM, typename ITERATOR>
inline std::vector<StreakBox> mergeStreaks(ITERATOR begin, const ITERATOR& end, const Coord<DIM>& gridOrigin)
{
    std::vector<StreakBox> boxes;
    /* rows -> slabs of rows (same x extent, consecutive y, one plane) */
    for (ITERATOR i = begin; i != end; ++i) {
        int32_t s[4];
        toStreak4(*i, gridOrigin, s);
        if (s[3] <= s[0]) {
            continue;
        }
        if (!boxes.empty()) {
            StreakBox& b = boxes.back();
            if (b.dim[2] == 1 && b.origin[0] == s[0] && b.dim[0] == s[3] - s[0] && b.origin[2] == s[2] && b.origin[1] + b.dim[1] == s[1]) {
                ++b.dim[1];
                continue;
            }
        }
        StreakBox b = {{s[0], s[1], s[2]}, {s[3] - s[0], 1, 1}};
        boxes.push_back(b);
    }
    /* slabs of rows -> boxes (same x and y extent, consecutive z) */
    std::size_t out = 0;
    for (std::size_t k = 0; k < boxes.size(); ++k) {
        if (out > 0) {
            StreakBox& b = boxes[out - 1];
            const StreakBox& n = boxes[k];
            if (b.origin[0] == n.origin[0] && b.dim[0] == n.dim[0] && b.origin[1] == n.origin[1] && b.dim[1] == n.dim[1] &&
                b.origin[2] + b.dim[2] == n.origin[2] && n.dim[2] == 1) {
                ++b.dim[2];
                continue;
            }
        }
        boxes[out++] = boxes[k];
    }
    boxes.resize(out);
    return boxes;
}
}

}

/* nvcc translation units: cells without a B200GEO_BIND_CELL line take the generic device path
 * (the user's own update() compiled into a kernel); with a host compiler unbound cells do not compile */
#include "b200genericsoa.h"
#ifdef __CUDACC__
#include "b200generic.h"
#endif

#include "b200boxgrid.h"
#include "b200containergrid.h"

namespace LibGeoDecomp {

template<typename CELL>
class B200Grid;

/* which device grid backs a cell type: the SoA grid, or the container grid for BoxCell<FixedArray<P, N> > */
template<typename CELL>
struct B200GridSelector {
    typedef B200Grid<CELL> Type;
};

template<typename PARTICLE, int N>
struct B200GridSelector<BoxCell<FixedArray<PARTICLE, N> > > {
    typedef B200BoxGrid<PARTICLE, N> Type;
};

/* ... or the ID-keyed container grid for ContainerCell<CARGO, SIZE> (the default int keys) */
template<typename CARGO, std::size_t SIZE>
struct B200GridSelector<ContainerCell<CARGO, SIZE, int> > {
    typedef B200ContainerGrid<CARGO, SIZE> Type;
};

template<typename CELL>
class B200Grid : public GridBase<CELL, APITraits::SelectTopology<CELL>::Value::DIM>
{
public:
    typedef typename APITraits::SelectTopology<CELL>::Value Topology;
    typedef typename APITraits::SelectStencil<CELL>::Value Stencil;
    static const int DIM = Topology::DIM;
    typedef GridBase<CELL, DIM> Base;

    explicit B200Grid(const CoordBox<DIM>& box = CoordBox<DIM>(), const CELL& edgeCell = CELL(), int device = 0) :
        Base(box.dimensions),
        box(box),
        edgeCell(edgeCell),
        device(device),
        handle(0),
        members(B200KernelBinding<CELL>::members()),
        slabGhost(0),
        lowPeer(false),
        highPeer(false)
    {
        init();
    }

    /* one slab of a larger simulation space (B200StripingSimulator): `slabGhost` ghost layers along the
     * last axis; a face towards a neighbouring slab holds that neighbour's cells (PEER) */
    B200Grid(const CoordBox<DIM>& box, const CELL& edgeCell, int device, int slabGhost, bool lowPeer, bool highPeer) :
        Base(box.dimensions),
        box(box),
        edgeCell(edgeCell),
        device(device),
        handle(0),
        members(B200KernelBinding<CELL>::members()),
        slabGhost(slabGhost),
        lowPeer(lowPeer),
        highPeer(highPeer)
    {
        init();
    }

    virtual ~B200Grid()
    {
        b200geo_grid_destroy(handle);
    }

    b200geo_grid *raw()
    {
        flush();
        return handle;
    }

    virtual void resize(const CoordBox<DIM>& newBox)
    {
        pendingCells.clear();
        pendingStreaks.clear();
        dropCaches();
        b200geo_grid_destroy(handle);
        handle = 0;
        box = newBox;
        this->topoDimensions = newBox.dimensions;
        create();
    }

    virtual void set(const Coord<DIM>& coord, const CELL& cell)
    {
        set(Streak<DIM>(coord, coord.x() + 1), &cell);
    }

    /* Writes are combined on the host and shipped as ONE streak list (b200geo_grid_load_region) before the
     * next read, sweep or explicit flush: an Initializer that calls set(Coord, cell) for every cell of its box
     * (src/examples/jacobi3d/main.cpp:54-72, src/examples/gameoflife/main.cpp:69-108) costs one transfer per
     * 2 MiB of cells instead of one per cell. Later writes win, as they would one by one. */
    virtual void set(const Streak<DIM>& streak, const CELL *cells)
    {
        int n = streak.length();
        if (n <= 0) {
            return;
        }
        int32_t s[4];
        B200Helpers::toStreak4(streak, box.origin, s);
        const std::size_t k = pendingStreaks.size();
        if (k >= 4 && pendingStreaks[k - 3] == s[1] && pendingStreaks[k - 2] == s[2] && pendingStreaks[k - 1] == s[0]) {
            pendingStreaks[k - 1] = s[3];   // continues the previous streak (cell after cell along x): one longer streak
        } else {
            pendingStreaks.insert(pendingStreaks.end(), s, s + 4);
        }
        if (n == 1) {
            pendingCells.push_back(*cells);
        } else {
            pendingCells.insert(pendingCells.end(), cells, cells + n);
        }
        dropCaches();
        if (pendingCells.size() >= maxPendingCells) {
            flush();
        }
    }

    /* ship the combined writes (both buffers: serialsimulator.h:54-57 initialises curGrid and newGrid alike).
     * The list is cut into runs of mutually disjoint streaks — one b200geo_grid_load_region call per run, in
     * order — so a cell that was written twice ends up with its last value (an Initializer that runs twice,
     * as in SerialSimulator's constructor + run(), gives two runs). */
    void flush() const
    {
        if (pendingCells.empty()) {
            return;
        }
        std::vector<CELL> cells;
        std::vector<int32_t> streaks;
        cells.swap(pendingCells);
        streaks.swap(pendingStreaks);
        pendingCells.reserve(cells.size());   // the next batch is likely as long: no regrowth from zero
        typedef std::vector<std::pair<int32_t, int32_t> > Intervals;
        std::map<std::pair<int32_t, int32_t>, Intervals> rows;
        std::size_t runStreak = 0, runCell = 0, cell = 0;
        // cell-by-cell Initializers stay in one row for a long time: look the row up once, not once per cell
        std::pair<int32_t, int32_t> lastRow(0, 0);
        Intervals *lastTaken = 0;
        for (std::size_t k = 0; k < streaks.size(); k += 4) {
            std::pair<int32_t, int32_t> row(streaks[k + 2], streaks[k + 1]);
            if (lastTaken == 0 || !(row == lastRow)) {
                lastTaken = &rows[row];
                lastRow = row;
            }
            Intervals& taken = *lastTaken;
            bool overlaps = false;
            for (std::size_t i = 0; i < taken.size(); ++i) {
                overlaps |= streaks[k] < taken[i].second && taken[i].first < streaks[k + 3];
            }
            if (overlaps) {
                ship(streaks, runStreak, k, cells, runCell, cell);
                rows.clear();
                runStreak = k;
                runCell = cell;
                lastTaken = &rows[row];
                lastTaken->push_back(std::make_pair(streaks[k], streaks[k + 3]));
            } else if (!taken.empty() && taken.back().second == streaks[k]) {
                taken.back().second = streaks[k + 3];   // cell after cell along x: one growing interval
            } else {
                taken.push_back(std::make_pair(streaks[k], streaks[k + 3]));
            }
            cell += streaks[k + 3] - streaks[k];
        }
        ship(streaks, runStreak, streaks.size(), cells, runCell, cell);
    }

    /* single-cell reads are served from a cache of the rows around them (Writers that walk the grid cell by
     * cell, e.g. TS_ASSERT_TEST_GRID, misc/testhelper.h:115-136): one transfer per block of up to 64 rows of a plane
     * (about 256 KiB), not one per cell */
    virtual CELL get(const Coord<DIM>& coord) const
    {
        bool hit = rowCacheValid;
        for (int d = 2; d < DIM && hit; ++d) {
            hit = coord[d] == rowCacheOrigin[d];
        }
        const int y = DIM > 1 ? coord[DIM > 1 ? 1 : 0] : 0, y0 = DIM > 1 ? rowCacheOrigin[DIM > 1 ? 1 : 0] : 0;
        hit = hit && y >= y0 && y < y0 + rowCacheRows && coord.x() >= box.origin.x() && coord.x() < box.origin.x() + box.dimensions.x();
        if (!hit) {
            bool inside = true;
            for (int d = 0; d < DIM; ++d) {
                inside &= coord[d] >= box.origin[d] && coord[d] < box.origin[d] + box.dimensions[d];
            }
            if (!inside) {
                // edge ring / ghost cells: straight from the device
                CELL cell;
                get(Streak<DIM>(coord, coord.x() + 1), &cell);
                return cell;
            }
            const int nx = box.dimensions.x();
            int rows = 1;
            if (DIM > 1) {
                int want = (int)((std::size_t)(1 << 18) / ((std::size_t)nx * sizeof(CELL)));
                rows = (std::max)(1, (std::min)((std::min)(want, 64), box.origin[DIM > 1 ? 1 : 0] + box.dimensions[DIM > 1 ? 1 : 0] - y));
            }
            Coord<DIM> origin = coord;
            origin.x() = box.origin.x();
            rowCache.resize((std::size_t)rows * nx);
            fetchRows(origin, rows, rowCache.data());
            rowCacheOrigin = origin;
            rowCacheRows = rows;
            rowCacheValid = true;
        }
        return rowCache[(std::size_t)(y - (DIM > 1 ? rowCacheOrigin[DIM > 1 ? 1 : 0] : 0)) * box.dimensions.x() + (coord.x() - box.origin.x())];
    }

    virtual void get(const Streak<DIM>& streak, CELL *cells) const
    {
        int n = streak.length();
        if (n <= 0) {
            return;
        }
        flush();
        int32_t s[4];
        B200Helpers::toStreak4(streak, box.origin, s);
        if (cellIsItsOnlyMember()) {
            B200Helpers::check(b200geo_grid_save_region(handle, s, 1, cells, B200GEO_HOST, 0));
            B200Helpers::check(b200geo_sync(0));
            return;
        }
        std::vector<char> buf((std::size_t)n * cellBytes);
        B200Helpers::check(b200geo_grid_save_region(handle, s, 1, buf.data(), B200GEO_HOST, 0));
        B200Helpers::check(b200geo_sync(0));
        std::size_t off = 0;
        for (std::size_t m = 0; m < members.size(); ++m) {
            for (int i = 0; i < n; ++i) {
                std::memcpy(reinterpret_cast<char*>(cells + i) + members[m].offsetInCell,
                            &buf[off + (std::size_t)i * members[m].bytes], members[m].bytes);
            }
            off += (std::size_t)n * members[m].bytes;
        }
    }

    virtual void setEdge(const CELL& cell)
    {
        edgeCell = cell;
        std::vector<char> buf(cellBytes);
        std::size_t off = 0;
        for (std::size_t m = 0; m < members.size(); ++m) {
            std::memcpy(&buf[off], reinterpret_cast<const char*>(&edgeCell) + members[m].offsetInCell, members[m].bytes);
            off += members[m].bytes;
        }
        B200Helpers::check(b200geo_grid_set_edge(handle, buf.data(), 0));
    }

    virtual const CELL& getEdge() const
    {
        return edgeCell;
    }

    virtual CoordBox<DIM> boundingBox() const
    {
        return box;
    }

    /* member-major byte stream, byte-compatible with SoAGrid::saveRegion (storage/soagrid.h:523-547) */
    virtual void saveRegion(std::vector<char> *buffer, const Region<DIM>& region, const Coord<DIM>& offset = Coord<DIM>()) const
    {
        flush();
        std::vector<int32_t> streaks = flatten(region, offset);
        buffer->resize(region.size() * cellBytes);
        B200Helpers::check(b200geo_grid_save_region(handle, streaks.data(), (int)(streaks.size() / 4), buffer->data(), B200GEO_HOST, 0));
        B200Helpers::check(b200geo_sync(0));
    }

    virtual void loadRegion(const std::vector<char>& buffer, const Region<DIM>& region, const Coord<DIM>& offset = Coord<DIM>())
    {
        if (buffer.size() != region.size() * cellBytes) {
            throw std::invalid_argument("buffer size does not match region");
        }
        flush();
        dropCaches();
        std::vector<int32_t> streaks = flatten(region, offset);
        B200Helpers::check(b200geo_grid_load_region(handle, streaks.data(), (int)(streaks.size() / 4), buffer.data(), B200GEO_HOST, 1, 0));
        B200Helpers::check(b200geo_sync(0));
    }

    /* the hot path: n sweeps + swaps on the device (SerialSimulator::nanoStep, serialsimulator.h:132-139) */
    void update(unsigned firstNanoStep, unsigned sweeps)
    {
        flush();
        dropCaches();
        int32_t dim[3] = {1, 1, 1};
        for (int i = 0; i < DIM; ++i) {
            dim[i] = box.dimensions[i];
        }
        B200KernelBinding<CELL>::step(handle, dim, firstNanoStep, sweeps);
    }

    void sync() const
    {
        flush();
        B200Helpers::check(b200geo_grid_sync(handle, 0));
    }

    /* UpdateFunctor over an arbitrary Region (storage/updatefunctor.h:403-428 takes a Region, too): one sweep of
     * the region's cells, current buffer -> scratch buffer, NO swap — one launch per box of the region
     * (B200Helpers::mergeStreaks). What a Stepper's update1() / updateGhost() do to inner sets and rims
     * (parallelization/nesting/vanillastepper.h:93-225). Returns the number of launches. */
    std::size_t updateRegion(const Region<DIM>& region, unsigned nanoStep)
    {
        flush();
        dropCaches();
        B200Helpers::check(b200geo_refresh_ghosts(handle, 0));
        std::vector<B200Helpers::StreakBox> boxes = B200Helpers::mergeStreaks<DIM>(region.beginStreak(), region.endStreak(), box.origin);
        for (std::size_t k = 0; k < boxes.size(); ++k) {
            B200KernelBinding<CELL>::updateBox(handle, nanoStep, boxes[k].origin, boxes[k].dim);
        }
        return boxes.size();
    }

    /* the scratch buffer becomes the current one (swap(oldGrid, newGrid)) */
    void swapBuffers()
    {
        flush();
        dropCaches();
        B200Helpers::check(b200geo_swap(handle));
    }

    /* region <-> a member-major buffer in DEVICE memory (b200geo_device_alloc): patch buffers that never leave the GPU */
    void saveRegionToDevice(void *deviceBuffer, const Region<DIM>& region) const
    {
        flush();
        std::vector<int32_t> streaks = flatten(region, Coord<DIM>());
        B200Helpers::check(b200geo_grid_save_region(handle, streaks.data(), (int)(streaks.size() / 4), deviceBuffer, B200GEO_CUDA_DEVICE, 0));
    }

    void loadRegionFromDevice(const void *deviceBuffer, const Region<DIM>& region)
    {
        flush();
        dropCaches();
        std::vector<int32_t> streaks = flatten(region, Coord<DIM>());
        /* the current buffer only: a Stepper's two grids differ on purpose */
        B200Helpers::check(b200geo_grid_load_region(handle, streaks.data(), (int)(streaks.size() / 4), deviceBuffer, B200GEO_CUDA_DEVICE, 0, 0));
    }

    /* GridBase::saveRegion / loadRegion for AoS buffers (std::vector<CELL>, storage/gridbase.h:157-180): what
     * SerializationBuffer selects for cells without an SoA registration (storage/serializationbuffer.h:20-57) */
    virtual void saveRegion(std::vector<CELL> *buffer, const Region<DIM>& region, const Coord<DIM>& offset = Coord<DIM>()) const
    {
        flush();
        std::vector<int32_t> streaks = flatten(region, offset);
        const std::size_t n = region.size();
        buffer->resize(n);
        if (n == 0) {
            return;
        }
        if (cellIsItsOnlyMember()) {
            B200Helpers::check(b200geo_grid_save_region(handle, streaks.data(), (int)(streaks.size() / 4), buffer->data(), B200GEO_HOST, 0));
            B200Helpers::check(b200geo_sync(0));
            return;
        }
        std::vector<char> raw(n * cellBytes);
        B200Helpers::check(b200geo_grid_save_region(handle, streaks.data(), (int)(streaks.size() / 4), raw.data(), B200GEO_HOST, 0));
        B200Helpers::check(b200geo_sync(0));
        std::size_t off = 0;
        for (std::size_t m = 0; m < members.size(); ++m) {
            const std::size_t bytes = members[m].bytes, at = members[m].offsetInCell;
            for (std::size_t i = 0; i < n; ++i) {
                std::memcpy(reinterpret_cast<char*>(&(*buffer)[i]) + at, &raw[off + i * bytes], bytes);
            }
            off += n * bytes;
        }
    }

    virtual void loadRegion(const std::vector<CELL>& buffer, const Region<DIM>& region, const Coord<DIM>& offset = Coord<DIM>())
    {
        loadRegionCells(buffer, region, offset, 1);
    }

    /* both = 0: the current buffer only */
    void loadRegionCells(const std::vector<CELL>& buffer, const Region<DIM>& region, const Coord<DIM>& offset, int both)
    {
        if (buffer.size() != region.size()) {
            throw std::invalid_argument("buffer size does not match region");
        }
        flush();
        dropCaches();
        const std::size_t n = region.size();
        if (n == 0) {
            return;
        }
        std::vector<int32_t> streaks = flatten(region, offset);
        if (cellIsItsOnlyMember()) {
            B200Helpers::check(b200geo_grid_load_region(handle, streaks.data(), (int)(streaks.size() / 4), buffer.data(), B200GEO_HOST, both, 0));
            B200Helpers::check(b200geo_sync(0));
            return;
        }
        std::vector<char> raw(n * cellBytes);
        std::size_t off = 0;
        for (std::size_t m = 0; m < members.size(); ++m) {
            const std::size_t bytes = members[m].bytes, at = members[m].offsetInCell;
            for (std::size_t i = 0; i < n; ++i) {
                std::memcpy(&raw[off + i * bytes], reinterpret_cast<const char*>(&buffer[i]) + at, bytes);
            }
            off += n * bytes;
        }
        B200Helpers::check(b200geo_grid_load_region(handle, streaks.data(), (int)(streaks.size() / 4), raw.data(), B200GEO_HOST, both, 0));
        B200Helpers::check(b200geo_sync(0));
    }

    /* the char-buffer twin of loadRegionCells */
    void loadRegionBytes(const std::vector<char>& buffer, const Region<DIM>& region, const Coord<DIM>& offset, int both)
    {
        if (buffer.size() != region.size() * cellBytes) {
            throw std::invalid_argument("buffer size does not match region");
        }
        flush();
        dropCaches();
        std::vector<int32_t> streaks = flatten(region, offset);
        B200Helpers::check(b200geo_grid_load_region(handle, streaks.data(), (int)(streaks.size() / 4), buffer.data(), B200GEO_HOST, both, 0));
        B200Helpers::check(b200geo_sync(0));
    }

    int deviceIndex() const
    {
        return device;
    }

    /* Member copies of plain selectors are enqueued without waiting for them (the caller synchronises): lets a slab
     * group keep the links of all its GPUs busy at once (B200StripedGrid). Page-locked host memory only — pageable
     * copies are synchronous anyway. */
    void setDeferSync(bool defer) const
    {
        deferSync = defer;
    }

    /* Selector I/O (loadMember / saveMember of plain member selectors) goes through `stream` from now on and is
     * not waited for; bothBuffers = false: a load writes the current buffer only, except for the members the
     * kernel family never rewrites. (0, true) restores the default: the null stream, both buffers, synchronous.
     * What B200GridWindow switches on around a plugin call of a streamed run. */
    void setMemberIo(void *stream, bool bothBuffers) const
    {
        ioStream = stream;
        ioBoth = bothBuffers;
    }

    /* combined host writes are shipped when this many cells are pending (default: 2 MiB worth, at least 64 Ki cells) */
    void setMaxPendingCells(std::size_t cells)
    {
        maxPendingCells = cells > 0 ? cells : 1;
    }

    /* the device grid was changed behind this object's back (slab group stepping): drop cached rows */
    void invalidateCache() const
    {
        dropCaches();
    }

    int bytesPerCell() const
    {
        return cellBytes;
    }

    /* strided member copies issued by saveMember / loadMember so far (one per box of the region) */
    std::size_t memberCopyCalls() const
    {
        return memberCalls;
    }

    /* Selector I/O for a streak list (what GridBase::saveMember / loadMember end up calling) */
    void saveMemberStreaks(
        char *target,
        MemoryLocation::Location targetLocation,
        const Selector<CELL>& selector,
        const typename Region<DIM>::StreakIterator& begin,
        const typename Region<DIM>::StreakIterator& end) const
    {
        saveMemberImplementation(target, targetLocation, selector, begin, end);
    }

    void loadMemberStreaks(
        const char *source,
        MemoryLocation::Location sourceLocation,
        const Selector<CELL>& selector,
        const typename Region<DIM>::StreakIterator& begin,
        const typename Region<DIM>::StreakIterator& end)
    {
        loadMemberImplementation(source, sourceLocation, selector, begin, end);
    }

protected:
    virtual void saveMemberImplementation(
        char *target,
        MemoryLocation::Location targetLocation,
        const Selector<CELL>& selector,
        const typename Region<DIM>::StreakIterator& begin,
        const typename Region<DIM>::StreakIterator& end) const
    {
        flush();
        int m = findMember(selector);
        if (m >= 0) {
            /* a plain member selector: one strided copy per BOX of the streak list, device to device for
             * MemoryLocation::CUDA_DEVICE (storage/gridbase.h:217-261) */
            const int loc = targetLocation == MemoryLocation::HOST ? B200GEO_HOST : B200GEO_CUDA_DEVICE;
            std::vector<B200Helpers::StreakBox> boxes = B200Helpers::mergeStreaks<DIM>(begin, end, box.origin);
            /* Writers that pull a member ROW BY ROW (the reference's BOVOutput::writeGrid calls saveMemberUnchecked
             * once per streak, io/bovoutput.h:83-95; so do PPM / VisIt writers): rows are served from a read-ahead
             * block of up to 64 MiB of the rows that follow — one transfer per block instead of one per row */
            if (ioStream == 0 && boxes.size() == 1 && boxes[0].dim[1] == 1 && boxes[0].dim[2] == 1 && loc == B200GEO_HOST &&
                boxes[0].origin[0] >= 0 && boxes[0].origin[0] + boxes[0].dim[0] <= box.dimensions.x() &&
                boxes[0].origin[1] >= 0 && boxes[0].origin[2] >= 0 && serveRow(m, boxes[0], target)) {
                return;
            }
            for (std::size_t k = 0; k < boxes.size(); ++k) {
                B200Helpers::check(b200geo_grid_save_member(handle, m, boxes[k].origin, boxes[k].dim, target, loc, ioStream));
                target += selector.sizeOfExternal() * boxes[k].cells();
            }
            memberCalls += boxes.size();
            if (!deferSync && ioStream == 0) {
                sync();
            }
            return;
        }
        for (typename Region<DIM>::StreakIterator i = begin; i != end; ++i) {
            int n = i->length();
            if (targetLocation != MemoryLocation::HOST) {
                throw std::logic_error("B200Grid: filtered selectors are only supported for host targets");
            }
            std::vector<CELL> cells(n);
            get(*i, cells.data());
            selector.copyMemberOut(cells.data(), MemoryLocation::HOST, target, MemoryLocation::HOST, n);
            target += selector.sizeOfExternal() * n;
        }
        sync();
    }

    virtual void loadMemberImplementation(
        const char *source,
        MemoryLocation::Location sourceLocation,
        const Selector<CELL>& selector,
        const typename Region<DIM>::StreakIterator& begin,
        const typename Region<DIM>::StreakIterator& end)
    {
        flush();
        dropCaches();
        int m = findMember(selector);
        if (m >= 0) {
            const int loc = sourceLocation == MemoryLocation::HOST ? B200GEO_HOST : B200GEO_CUDA_DEVICE;
            std::vector<B200Helpers::StreakBox> boxes = B200Helpers::mergeStreaks<DIM>(begin, end, box.origin);
            for (std::size_t k = 0; k < boxes.size(); ++k) {
                const int both = (ioBoth || B200Helpers::invariantMember(B200KernelBinding<CELL>::kernel(), m)) ? 1 : 0;
                B200Helpers::check(b200geo_grid_load_member(handle, m, boxes[k].origin, boxes[k].dim, source, loc, both, ioStream));
                source += selector.sizeOfExternal() * boxes[k].cells();
            }
            memberCalls += boxes.size();
            if (!deferSync && ioStream == 0) {
                sync();
            }
            return;
        }
        for (typename Region<DIM>::StreakIterator i = begin; i != end; ++i) {
            int n = i->length();
            if (sourceLocation != MemoryLocation::HOST) {
                throw std::logic_error("B200Grid: filtered selectors are only supported for host sources");
            }
            std::vector<CELL> cells(n);
            get(*i, cells.data());
            selector.copyMemberIn(source, MemoryLocation::HOST, cells.data(), MemoryLocation::HOST, n);
            set(*i, cells.data());
            source += selector.sizeOfExternal() * n;
        }
        sync();
    }

private:
    CoordBox<DIM> box;
    CELL edgeCell;
    int device;
    b200geo_grid *handle;
    std::vector<B200Member> members;
    int cellBytes;
    int slabGhost;
    bool lowPeer;
    bool highPeer;
    std::size_t maxPendingCells = ((std::size_t)2 << 20) / sizeof(CELL) > (1 << 16) ? ((std::size_t)2 << 20) / sizeof(CELL) : (1 << 16);   /* 2 MiB of combined writes per flush (still cache resident when the copy reads them; 32 MiB measured 1.6 x slower) */
    mutable std::vector<CELL> pendingCells;        /* combined writes: cells ...                  */
    mutable std::vector<int32_t> pendingStreaks;   /* ... and where they go, {x, y, z, endX} each */
    mutable std::vector<CELL> rowCache;            /* rowCacheRows whole rows of one plane, starting at rowCacheOrigin */
    mutable Coord<DIM> rowCacheOrigin;
    mutable int rowCacheRows = 0;
    mutable bool rowCacheValid = false;
    mutable std::size_t memberCalls = 0;
    mutable bool deferSync = false;
    mutable void *ioStream = 0;                    /* setMemberIo */
    mutable bool ioBoth = true;
    mutable std::string lastSelectorName;          /* findMember: the selector asked for last and its member */
    mutable std::size_t lastSelectorBytes = 0;
    mutable int lastSelectorMember = -1;
    mutable std::vector<char> memberCache;         /* read-ahead block of one member (saveMember row by row) */
    mutable B200Helpers::StreakBox memberCacheBox;
    mutable int memberCacheMember = -1;
    mutable bool memberCacheValid = false;

    void dropCaches() const
    {
        rowCacheValid = false;
        memberCacheValid = false;
    }

    /* one row of member m out of the read-ahead block; the block is (re)filled with the rows from this one on:
     * the rest of the plane, or — from the first row of a plane — as many whole planes as fit */
    bool serveRow(int m, const B200Helpers::StreakBox& row, char *target) const
    {
        const int nx = box.dimensions.x();
        const int ny = DIM > 1 ? box.dimensions[DIM > 1 ? 1 : 0] : 1;
        const int nz = DIM > 2 ? box.dimensions[DIM > 2 ? 2 : 0] : 1;
        const std::size_t bytes = members[m].bytes;
        const int y = row.origin[1], z = row.origin[2];
        if (y >= ny || z >= nz) {
            return false;
        }
        bool hit = memberCacheValid && memberCacheMember == m && z >= memberCacheBox.origin[2] &&
                   z < memberCacheBox.origin[2] + memberCacheBox.dim[2] && y >= memberCacheBox.origin[1] &&
                   y < memberCacheBox.origin[1] + memberCacheBox.dim[1];
        if (!hit) {
            const std::size_t budget = (std::size_t)64 << 20, rowBytes = (std::size_t)nx * bytes;
            B200Helpers::StreakBox b = {{0, y, z}, {nx, 1, 1}};
            if (y == 0 && rowBytes * ny <= budget) {
                b.dim[1] = ny;
                b.dim[2] = (int)(std::min)((std::size_t)(nz - z), budget / (rowBytes * ny));
            } else {
                b.dim[1] = (int)(std::max)((std::size_t)1, (std::min)((std::size_t)(ny - y), budget / rowBytes));
            }
            if (b.cells() < 2 * (std::size_t)nx) {
                return false;      /* nothing to read ahead: the plain path */
            }
            memberCache.resize(b.cells() * bytes);
            B200Helpers::check(b200geo_grid_save_member(handle, m, b.origin, b.dim, memberCache.data(), B200GEO_HOST, 0));
            B200Helpers::check(b200geo_sync(0));
            ++memberCalls;
            memberCacheBox = b;
            memberCacheMember = m;
            memberCacheValid = true;
        }
        const std::size_t at = (((std::size_t)(z - memberCacheBox.origin[2]) * memberCacheBox.dim[1] + (y - memberCacheBox.origin[1])) * nx +
                                row.origin[0]) * bytes;
        std::memcpy(target, &memberCache[at], (std::size_t)row.dim[0] * bytes);
        return true;
    }

    /* `rows` whole rows (same plane, consecutive y) starting at `origin` in ONE transfer */
    void fetchRows(const Coord<DIM>& origin, int rows, CELL *cells) const
    {
        flush();
        const int nx = box.dimensions.x();
        std::vector<int32_t> streaks((std::size_t)rows * 4);
        for (int r = 0; r < rows; ++r) {
            Coord<DIM> c = origin;
            if (DIM > 1) {
                c[DIM > 1 ? 1 : 0] += r;
            }
            B200Helpers::toStreak4(Streak<DIM>(c, c.x() + nx), box.origin, &streaks[(std::size_t)r * 4]);
        }
        const std::size_t n = (std::size_t)rows * nx;
        if (cellIsItsOnlyMember()) {
            B200Helpers::check(b200geo_grid_save_region(handle, streaks.data(), rows, cells, B200GEO_HOST, 0));
            B200Helpers::check(b200geo_sync(0));
            return;
        }
        std::vector<char> buf(n * cellBytes);
        B200Helpers::check(b200geo_grid_save_region(handle, streaks.data(), rows, buf.data(), B200GEO_HOST, 0));
        B200Helpers::check(b200geo_sync(0));
        std::size_t off = 0;
        for (std::size_t m = 0; m < members.size(); ++m) {
            const std::size_t bytes = members[m].bytes, at = members[m].offsetInCell;
            for (std::size_t i = 0; i < n; ++i) {
                std::memcpy(reinterpret_cast<char*>(cells + i) + at, &buf[off + i * bytes], bytes);
            }
            off += n * bytes;
        }
    }

    /* streaks [s0, s1) of the list with their cells [c0, c1), member-major, in one call */
    void ship(const std::vector<int32_t>& streaks, std::size_t s0, std::size_t s1,
              const std::vector<CELL>& cells, std::size_t c0, std::size_t c1) const
    {
        if (c1 <= c0) {
            return;
        }
        std::size_t n = c1 - c0;
        if (cellIsItsOnlyMember()) {
            // an array of such cells IS the member-major stream (Jacobi, Game of Life): no packing pass
            B200Helpers::check(b200geo_grid_load_region(handle, &streaks[s0], (int)((s1 - s0) / 4), &cells[c0], B200GEO_HOST, 1, 0));
            return;
        }
        std::vector<char> buf(n * cellBytes);
        std::size_t off = 0;
        for (std::size_t m = 0; m < members.size(); ++m) {
            const std::size_t bytes = members[m].bytes, at = members[m].offsetInCell;
            for (std::size_t i = 0; i < n; ++i) {
                std::memcpy(&buf[off + i * bytes], reinterpret_cast<const char*>(&cells[c0 + i]) + at, bytes);
            }
            off += n * bytes;
        }
        B200Helpers::check(b200geo_grid_load_region(handle, &streaks[s0], (int)((s1 - s0) / 4), buf.data(), B200GEO_HOST, 1, 0));
    }

    bool cellIsItsOnlyMember() const
    {
        return members.size() == 1 && members[0].offsetInCell == 0 && (std::size_t)members[0].bytes == sizeof(CELL);
    }

    void init()
    {
        cellBytes = 0;
        for (std::size_t m = 0; m < members.size(); ++m) {
            cellBytes += members[m].bytes;
        }
        create();
    }

    void create()
    {
        b200geo_grid_desc desc;
        std::memset(&desc, 0, sizeof(desc));
        for (int i = 0; i < 3; ++i) {
            bool used = i < DIM;
            desc.dim[i] = used ? box.dimensions[i] : 1;
            desc.ghost[i] = used ? Stencil::RADIUS : 0;
            int mode = (used && Topology::wrapsAxis(i)) ? B200GEO_GHOST_WRAP : B200GEO_GHOST_EDGE;
            desc.ghost_mode[i][0] = desc.ghost_mode[i][1] = mode;
            // periodic images 4 cells wide let the temporal-blocked Jacobi kernels fuse up to 4 sweeps
            // per launch on a Torus (a wrap ghost may not be wider than the grid itself)
            int k = B200KernelBinding<CELL>::kernel();
            bool jacobi = k == B200GEO_KERNEL_JACOBI6 || k == B200GEO_KERNEL_JACOBI7 || k == B200GEO_KERNEL_JACOBI27;
            if (jacobi && mode == B200GEO_GHOST_WRAP) {
                desc.ghost[i] = std::max(desc.ghost[i], std::min(4, desc.dim[i]));
            }
            if (i == DIM - 1 && slabGhost > 0) {
                // a slab: the faces towards its neighbours are filled by the halo exchange; an outer face
                // of a Torus slab is PEER as well (ring closure), so only Cube slabs keep EDGE here
                desc.ghost[i] = slabGhost;
                desc.ghost_mode[i][0] = lowPeer ? B200GEO_GHOST_PEER : B200GEO_GHOST_EDGE;
                desc.ghost_mode[i][1] = highPeer ? B200GEO_GHOST_PEER : B200GEO_GHOST_EDGE;
            }
        }
        if (members.size() > B200GEO_MAX_MEMBERS) {
            throw std::out_of_range("too many SoA members");
        }
        desc.n_members = (int)members.size();
        for (std::size_t m = 0; m < members.size(); ++m) {
            desc.member_bytes[m] = members[m].bytes;
        }
        B200Helpers::check(B200Helpers::createGrid<B200KernelBinding<CELL> >(&desc, device, &handle, 0));
        setEdge(edgeCell);
    }

    std::vector<int32_t> flatten(const Region<DIM>& region, const Coord<DIM>& offset) const
    {
        std::vector<int32_t> streaks;
        streaks.reserve(region.numStreaks() * 4);
        for (typename Region<DIM>::StreakIterator i = region.beginStreak(); i != region.endStreak(); ++i) {
            Streak<DIM> s = *i;
            s.origin += offset;
            s.endX += offset.x();
            int32_t v[4];
            B200Helpers::toStreak4(s, box.origin, v);
            streaks.insert(streaks.end(), v, v + 4);
        }
        return streaks;
    }

    /* plain pointer-to-member selectors of a bound member map straight onto one device array */
    int findMember(const Selector<CELL>& selector) const
    {
        if (selector.sizeOfExternal() != selector.sizeOfMember() || selector.arity() != 1) {
            return -1;
        }
        /* Writers ask again and again with the same selector (row by row): remember the last answer */
        if (lastSelectorMember >= 0 && selector.name() == lastSelectorName && selector.sizeOfMember() == lastSelectorBytes) {
            return lastSelectorMember;
        }
        int found = probeMember(selector);
        if (found >= 0) {
            lastSelectorName = selector.name();
            lastSelectorBytes = selector.sizeOfMember();
            lastSelectorMember = found;
        }
        return found;
    }

    int probeMember(const Selector<CELL>& selector) const
    {
        CELL probe = CELL();
        for (std::size_t m = 0; m < members.size(); ++m) {
            if ((std::size_t)members[m].bytes != selector.sizeOfMember()) {
                continue;
            }
            /* identify the member by round-tripping a byte pattern through the selector */
            CELL c = probe;
            std::vector<char> pattern(members[m].bytes);
            for (int b = 0; b < members[m].bytes; ++b) {
                pattern[b] = (char)(0x5a + b);
            }
            std::memcpy(reinterpret_cast<char*>(&c) + members[m].offsetInCell, pattern.data(), members[m].bytes);
            std::vector<char> out(members[m].bytes);
            selector.copyMemberOut(&c, MemoryLocation::HOST, out.data(), MemoryLocation::HOST, 1);
            if (std::memcmp(out.data(), pattern.data(), members[m].bytes) == 0) {
                return (int)m;
            }
        }
        return -1;
    }
};

}

#include "b200streamedrun.h"

namespace LibGeoDecomp {

template<typename CELL>
class B200Simulator : public MonolithicSimulator<CELL>
{
public:
    typedef typename MonolithicSimulator<CELL>::Topology Topology;
    typedef typename MonolithicSimulator<CELL>::WriterVector WriterVector;
    typedef typename Steerer<CELL>::SteererFeedback SteererFeedback;
    typedef typename B200GridSelector<CELL>::Type GridType;
    typedef GridBase<CELL, Topology::DIM> GridBaseType;
    static const int DIM = Topology::DIM;
    static const unsigned NANO_STEPS = APITraits::SelectNanoSteps<CELL>::VALUE;

    using MonolithicSimulator<CELL>::chronometer;
    using MonolithicSimulator<CELL>::getStep;
    using MonolithicSimulator<CELL>::initializer;
    using MonolithicSimulator<CELL>::gridDim;
    using MonolithicSimulator<CELL>::steerers;
    using MonolithicSimulator<CELL>::stepNum;
    using MonolithicSimulator<CELL>::writers;

    explicit B200Simulator(Initializer<CELL> *init, int device = 0) :
        MonolithicSimulator<CELL>(init),
        grid(CoordBox<DIM>(Coord<DIM>(), init->gridBox().dimensions), CELL(), device)
    {
        stepNum = init->startStep();
        simArea << CoordBox<DIM>(Coord<DIM>(), init->gridBox().dimensions);
        // both device buffers receive the initial state (serialsimulator.h:54-57)
        initializer->grid(&grid);
    }

    /* Call sites written for CUDASimulator(initializer, blockSize) (parallelization/cudasimulator.h:314-316) and for
     * OpenMPSimulator(initializer, enableFineGrainedParallelism) (openmpsimulator.h:48-50) keep compiling and keep their
     * meaning: the launch geometry is the kernels' own business here (b200geo_set_tuning), and a bool is NOT a device id. */
    B200Simulator(Initializer<CELL> *init, const Coord<3>& /* blockSize */, int device = 0) :
        MonolithicSimulator<CELL>(init),
        grid(CoordBox<DIM>(Coord<DIM>(), init->gridBox().dimensions), CELL(), device)
    {
        stepNum = init->startStep();
        simArea << CoordBox<DIM>(Coord<DIM>(), init->gridBox().dimensions);
        initializer->grid(&grid);
    }

    B200Simulator(Initializer<CELL> *init, bool /* enableFineGrainedParallelism */) :
        MonolithicSimulator<CELL>(init),
        grid(CoordBox<DIM>(Coord<DIM>(), init->gridBox().dimensions), CELL(), 0)
    {
        stepNum = init->startStep();
        simArea << CoordBox<DIM>(Coord<DIM>(), init->gridBox().dimensions);
        initializer->grid(&grid);
    }

    using MonolithicSimulator<CELL>::addWriter;

    /* ParallelWriters (io/parallelwriter.h; what programs written for the reference's StripingSimulator /
     * HiParSimulator register): called with the whole area, rank 0 and lastCall = true by a plain run — and chunk by
     * chunk, while the sweeps go on, by a streamed one (b200streamedrun.h). The simulator takes ownership, like the
     * reference. Writers that are BOTH a Writer and a ParallelWriter keep going through addWriter(Writer *). */
    template<typename WRITER>
    typename std::enable_if<std::is_base_of<ParallelWriter<CELL>, WRITER>::value &&
                            !std::is_base_of<Writer<CELL>, WRITER>::value>::type
    addWriter(WRITER *writer)
    {
        parallelWriters.push_back(typename SharedPtr<ParallelWriter<CELL> >::Type(writer));
    }

    /* one step, exactly like the reference (serialsimulator.h:70-93) */
    virtual void step()
    {
        SteererFeedback feedback;
        step(&feedback, false);
    }

    virtual void run()
    {
        for (std::size_t i = 0; i < parallelWriters.size(); ++i) {
            parallelWriters[i]->setRegion(simArea);
        }
        if (streamIO && steerers.empty() && writers.empty() && runStreamed(&grid)) {
            return;
        }
        initializer->grid(&grid);
        stepNum = initializer->startStep();
        for (unsigned i = 0; i < steerers.size(); i++) {
            steerers[i]->setRegion(simArea);
        }

        SteererFeedback feedback;
        handleInput(STEERER_INITIALIZED, &feedback);
        handleOutput(WRITER_INITIALIZED);

        for (; stepNum < initializer->maxSteps();) {
            if (feedback.simulationEnded()) {
                break;
            }
            step(&feedback, fuseSteps);
        }

        handleInput(STEERER_ALL_DONE, &feedback);
        grid.sync();
    }

    virtual const GridBaseType *getGrid()
    {
        grid.sync();
        return &grid;
    }

    /* run() pipelines Initializer, sweeps and ParallelWriters chunk by chunk (b200streamedrun.h) when nothing but
     * ParallelWriters with no call due inside the run is registered, the cell is bound to a kernel family that updates
     * boxes, and the last axis does not wrap; false: always upload, sweep, download one after the other */
    bool streamIO = true;
    int streamChunks = 16;

    /* how many run() calls took the streamed schedule */
    std::size_t streamedRuns() const
    {
        return streamed;
    }

protected:
    GridType grid;
    Region<DIM> simArea;
    std::vector<typename SharedPtr<ParallelWriter<CELL> >::Type> parallelWriters;
    std::size_t streamed = 0;

    bool runStreamed(B200Grid<CELL> *target)
    {
        B200StreamedRun<CELL> schedule;
        if (!schedule.plan(target->boundingBox(), initializer->startStep(), initializer->maxSteps(), parallelWriters, streamChunks)) {
            return false;
        }
        TimeTotal t(&chronometer);
        stepNum = initializer->startStep();
        schedule.run(target, &*initializer, parallelWriters, gridDim);
        stepNum = initializer->maxSteps();
        ++streamed;
        return true;
    }

    /* container grids (n-body) have no streamed schedule */
    template<typename OTHER_GRID>
    bool runStreamed(OTHER_GRID *)
    {
        return false;
    }

    void step(SteererFeedback *feedback, bool fuse)
    {
        TimeTotal t(&chronometer);
        handleInput(STEERER_NEXT_STEP, feedback);

        // inside run(): fuse every step up to the next observable event (writer / steerer period,
        // maxSteps) into one engine call; with no plugins due this is the whole remaining run
        unsigned steps = fuse ? stepsToNextEvent() : 1;
        {
            TimeCompute t(&chronometer);
            grid.update(0, steps * NANO_STEPS);
            if (steps > 1 || !writers.empty() || !parallelWriters.empty()) {
                grid.sync();
            }
        }
        stepNum += steps;

        WriterEvent event = WRITER_STEP_FINISHED;
        if (stepNum == initializer->maxSteps()) {
            event = WRITER_ALL_DONE;
        }
        handleOutput(event);
    }

    unsigned stepsToNextEvent() const
    {
        unsigned max = initializer->maxSteps();
        unsigned n = (stepNum < max) ? (max - stepNum) : 1;
        for (unsigned i = 0; i < writers.size(); ++i) {
            unsigned p = writers[i]->getPeriod();
            n = (std::min)(n, p - stepNum % p);
        }
        for (unsigned i = 0; i < steerers.size(); ++i) {
            unsigned p = steerers[i]->getPeriod();
            n = (std::min)(n, p - stepNum % p);
        }
        for (std::size_t i = 0; i < parallelWriters.size(); ++i) {
            unsigned p = parallelWriters[i]->getPeriod();
            n = (std::min)(n, p - stepNum % p);
        }
        return n > 0 ? n : 1;
    }

    void handleOutput(WriterEvent event)
    {
        TimeOutput t(&chronometer);
        for (unsigned i = 0; i < writers.size(); i++) {
            if ((event != WRITER_STEP_FINISHED) || ((getStep() % writers[i]->getPeriod()) == 0)) {
                writers[i]->stepFinished(grid, getStep(), event);
            }
        }
        for (std::size_t i = 0; i < parallelWriters.size(); ++i) {
            if ((event != WRITER_STEP_FINISHED) || ((getStep() % parallelWriters[i]->getPeriod()) == 0)) {
                grid.sync();
                parallelWriters[i]->stepFinished(grid, simArea, gridDim, getStep(), event, 0, true);
            }
        }
    }

    void handleInput(SteererEvent event, SteererFeedback *feedback)
    {
        TimeInput t(&chronometer);
        for (unsigned i = 0; i < steerers.size(); ++i) {
            if ((event != STEERER_NEXT_STEP) || (stepNum % steerers[i]->getPeriod() == 0)) {
                steerers[i]->nextStep(&grid, simArea, gridDim, getStep(), event, 0, true, feedback);
            }
        }
    }

public:
    /* run() fuses the steps between two plugin events into one engine call (no observable
     * difference: plugins see the same (step, event) sequence); set to false for one call per step. */
    bool fuseSteps = true;
};

}

#endif
