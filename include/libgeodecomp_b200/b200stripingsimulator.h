/* Multi-GPU drop-in: ONE process and host thread, the simulation space cut into equal
 * slabs along the last axis (geometry/partitions/stripingpartition.h:57-62), one slab per GPU of the box.
 *
 *   B200StripedGrid<CELL>        the GridBase<CELL, DIM> that Initializers, Writers and Steerers see: the
 *                                whole simulation space; every access is routed to the slab that owns it.
 *   B200StripingSimulator<CELL>  event protocol of SerialSimulator (parallelization/serialsimulator.h:48-187),
 *                                partition and schedule of StripingSimulator
 *                                (parallelization/stripingsimulator.h:269-286: rims first, ship them, interior
 *                                while they travel) with ghost zone width k as in HiParSimulator /
 *                                VanillaStepper (parallelization/nesting/vanillastepper.h:157-225) — but the
 *                                PatchLink MPI messages (communication/patchlink.h:127-151, 218-244) are
 *                                direct NVLink copies between the GPUs' ghost planes (b200geo_group_*).
 *
 * The reference needs one MPI rank per core/GPU for this; here the user's main() stays the serial one:
 *     B200StripingSimulator<Cell> sim(new MyInitializer(...));          // all GPUs of the box
 *     sim.addWriter(...); sim.run();
 * Results are bit-identical to SerialSimulator for every slab count (tests/facade/striping_test.cpp).
 */
#ifndef LIBGEODECOMP_B200_B200STRIPINGSIMULATOR_H
#define LIBGEODECOMP_B200_B200STRIPINGSIMULATOR_H

#include "b200simulator.h"

#include <libgeodecomp/io/parallelwriter.h>
#include <libgeodecomp/loadbalancer/loadbalancer.h>
#include <libgeodecomp/misc/sharedptr.h>
#include <libgeodecomp/storage/selector.h>

#include <memory>
#include <type_traits>

namespace LibGeoDecomp {

template<typename CELL>
class B200StripedGrid : public GridBase<CELL, APITraits::SelectTopology<CELL>::Value::DIM>
{
public:
    typedef typename APITraits::SelectTopology<CELL>::Value Topology;
    static const int DIM = Topology::DIM;
    static const int LAST = DIM - 1;
    typedef GridBase<CELL, DIM> Base;
    typedef B200Grid<CELL> SlabType;

    /* devices[s] = CUDA device of slab s (devices may repeat: several slabs on one GPU) */
    B200StripedGrid(const CoordBox<DIM>& box, const std::vector<int>& devices, int ghostWidth, const CELL& edgeCell = CELL()) :
        Base(box.dimensions),
        box(box),
        edgeCell(edgeCell),
        group(0),
        members(B200KernelBinding<CELL>::members())
    {
        int n = (int)devices.size();
        if (n < 1) {
            throw std::invalid_argument("B200StripedGrid needs at least one device");
        }
        int extent = box.dimensions[LAST];
        if (n > 1 && extent / n < ghostWidth) {
            throw std::invalid_argument("slab thinner than the ghost zone");
        }
        bool periodic = Topology::wrapsAxis(LAST);
        std::vector<b200geo_grid*> handles;
        for (int s = 0; s <= n; ++s) {
            bounds.push_back(box.origin[LAST] + (int)(((long)extent * s) / n));
        }
        for (int s = 0; s < n; ++s) {
            CoordBox<DIM> slabBox = box;
            slabBox.origin[LAST] = bounds[s];
            slabBox.dimensions[LAST] = bounds[s + 1] - bounds[s];
            if (n == 1) {
                slabs.push_back(std::unique_ptr<SlabType>(new SlabType(slabBox, edgeCell, devices[s])));
            } else {
                bool low = s > 0 || periodic, high = s < n - 1 || periodic;
                slabs.push_back(std::unique_ptr<SlabType>(new SlabType(slabBox, edgeCell, devices[s], ghostWidth, low, high)));
            }
            handles.push_back(slabs.back()->raw());
        }
        B200Helpers::check(b200geo_group_create(handles.data(), n, periodic && n > 1, &group));
        cellBytes = slabs[0]->bytesPerCell();
    }

    virtual ~B200StripedGrid()
    {
        b200geo_group_destroy(group);
    }

    /* ghost zone width = sweeps between two halo exchanges; the kernel families that fuse sweeps (temporal-blocked
     * Jacobi, two-sweep LBM) take the sweeps of one round in a single launch per rim / interior, so 2 is their default */
    static int defaultGhostWidth()
    {
        int radius = APITraits::SelectStencil<CELL>::Value::RADIUS;
        return std::max(radius, B200Helpers::fusedSweeps(B200KernelBinding<CELL>::kernel()) >= 2 ? 2 : 1);
    }

    std::size_t numSlabs() const
    {
        return slabs.size();
    }

    const SlabType& slab(std::size_t s) const
    {
        return *slabs[s];
    }

    /* plane range [first, second) of slab s along the last axis */
    std::pair<int, int> slabRange(std::size_t s) const
    {
        return std::make_pair(bounds[s], bounds[s + 1]);
    }

    virtual void resize(const CoordBox<DIM>&)
    {
        throw std::logic_error("B200StripedGrid cannot be resized");
    }

    virtual void set(const Coord<DIM>& coord, const CELL& cell)
    {
        slabs[owner(coord[LAST])]->set(coord, cell);
        dirty = true;
    }

    virtual void set(const Streak<DIM>& streak, const CELL *cells)
    {
        slabs[owner(streak.origin[LAST])]->set(streak, cells);
        dirty = true;
    }

    virtual CELL get(const Coord<DIM>& coord) const
    {
        return slabs[owner(coord[LAST])]->get(coord);
    }

    virtual void get(const Streak<DIM>& streak, CELL *cells) const
    {
        slabs[owner(streak.origin[LAST])]->get(streak, cells);
    }

    virtual void setEdge(const CELL& cell)
    {
        edgeCell = cell;
        for (std::size_t s = 0; s < slabs.size(); ++s) {
            slabs[s]->setEdge(cell);
        }
    }

    virtual const CELL& getEdge() const
    {
        return edgeCell;
    }

    virtual CoordBox<DIM> boundingBox() const
    {
        return box;
    }

    /* member-major like SoAGrid::saveRegion (storage/soagrid.h:523-547). A Region orders its streaks by
     * the last axis first, so the cells of one slab are ONE contiguous run of every member's block. */
    virtual void saveRegion(std::vector<char> *buffer, const Region<DIM>& region, const Coord<DIM>& offset = Coord<DIM>()) const
    {
        buffer->resize(region.size() * cellBytes);
        std::vector<Region<DIM> > parts = split(region, offset);
        std::size_t before = 0;
        for (std::size_t s = 0; s < slabs.size(); ++s) {
            std::size_t count = parts[s].size();
            if (count == 0) {
                continue;
            }
            std::vector<char> chunk;
            slabs[s]->saveRegion(&chunk, parts[s]);
            std::size_t memberOffset = 0;
            for (std::size_t m = 0; m < members.size(); ++m) {
                std::size_t b = members[m].bytes;
                std::memcpy(buffer->data() + memberOffset * region.size() + before * b, chunk.data() + memberOffset * count, count * b);
                memberOffset += b;
            }
            before += count;
        }
    }

    virtual void loadRegion(const std::vector<char>& buffer, const Region<DIM>& region, const Coord<DIM>& offset = Coord<DIM>())
    {
        if (buffer.size() != region.size() * cellBytes) {
            throw std::invalid_argument("buffer size does not match region");
        }
        std::vector<Region<DIM> > parts = split(region, offset);
        std::size_t before = 0;
        for (std::size_t s = 0; s < slabs.size(); ++s) {
            std::size_t count = parts[s].size();
            if (count == 0) {
                continue;
            }
            std::vector<char> chunk(count * cellBytes);
            std::size_t memberOffset = 0;
            for (std::size_t m = 0; m < members.size(); ++m) {
                std::size_t b = members[m].bytes;
                std::memcpy(chunk.data() + memberOffset * count, buffer.data() + memberOffset * region.size() + before * b, count * b);
                memberOffset += b;
            }
            slabs[s]->loadRegion(chunk, parts[s]);
            before += count;
        }
        dirty = true;
    }

    /* the AoS-buffer variants (storage/gridbase.h:157-180): streaks are ordered by the last axis first, so every
     * slab's cells are one contiguous block of the buffer */
    virtual void saveRegion(std::vector<CELL> *buffer, const Region<DIM>& region, const Coord<DIM>& offset = Coord<DIM>()) const
    {
        buffer->resize(region.size());
        std::vector<Region<DIM> > parts = split(region, offset);
        std::size_t before = 0;
        std::vector<CELL> chunk;
        for (std::size_t s = 0; s < slabs.size(); ++s) {
            if (parts[s].size() == 0) {
                continue;
            }
            slabs[s]->saveRegion(&chunk, parts[s]);
            std::copy(chunk.begin(), chunk.end(), buffer->begin() + before);
            before += chunk.size();
        }
    }

    virtual void loadRegion(const std::vector<CELL>& buffer, const Region<DIM>& region, const Coord<DIM>& offset = Coord<DIM>())
    {
        if (buffer.size() != region.size()) {
            throw std::invalid_argument("buffer size does not match region");
        }
        std::vector<Region<DIM> > parts = split(region, offset);
        std::size_t before = 0;
        for (std::size_t s = 0; s < slabs.size(); ++s) {
            std::size_t count = parts[s].size();
            if (count == 0) {
                continue;
            }
            std::vector<CELL> chunk(buffer.begin() + before, buffer.begin() + before + count);
            slabs[s]->loadRegion(chunk, parts[s]);
            before += count;
        }
        dirty = true;
    }

    /* sweeps x { update every slab; swap } with the halo exchanges they need */
    void update(unsigned firstNanoStep, unsigned sweeps)
    {
        for (std::size_t s = 0; s < slabs.size(); ++s) {
            slabs[s]->flush();            // combined host writes reach the device before the sweep
            slabs[s]->invalidateCache();  // and cached rows are stale after it
        }
        if (dirty) {
            // cells were written from the host: the neighbours' ghost copies are stale
            B200Helpers::check(b200geo_group_invalidate(group));
            dirty = false;
        }
        // bound cells: the library's kernels; unbound cells (nvcc translation units): their own update()
        B200KernelBinding<CELL>::groupStep(group, firstNanoStep, sweeps);
    }

    void sync() const
    {
        for (std::size_t s = 0; s < slabs.size(); ++s) {
            slabs[s]->flush();
        }
        B200Helpers::check(b200geo_group_sync(group));
    }

    /* halo traffic so far: (number of exchanges, bytes shipped between slabs) */
    std::pair<unsigned long long, unsigned long long> exchangeStatistics() const
    {
        uint64_t out[2] = {0, 0};
        B200Helpers::check(b200geo_group_stats(group, out));
        return std::make_pair((unsigned long long)out[0], (unsigned long long)out[1]);
    }

protected:
    virtual void saveMemberImplementation(
        char *target,
        MemoryLocation::Location targetLocation,
        const Selector<CELL>& selector,
        const typename Region<DIM>::StreakIterator& begin,
        const typename Region<DIM>::StreakIterator& end) const
    {
        /* all slabs' copies are enqueued first and awaited afterwards: the GPUs' links work at the same time */
        std::vector<Region<DIM> > parts = split(begin, end);
        for (std::size_t s = 0; s < slabs.size(); ++s) {
            if (parts[s].size() == 0) {
                continue;
            }
            slabs[s]->setDeferSync(true);
            slabs[s]->saveMemberStreaks(target, targetLocation, selector, parts[s].beginStreak(), parts[s].endStreak());
            slabs[s]->setDeferSync(false);
            target += selector.sizeOfExternal() * parts[s].size();
        }
        for (std::size_t s = 0; s < slabs.size(); ++s) {
            slabs[s]->sync();
        }
    }

    virtual void loadMemberImplementation(
        const char *source,
        MemoryLocation::Location sourceLocation,
        const Selector<CELL>& selector,
        const typename Region<DIM>::StreakIterator& begin,
        const typename Region<DIM>::StreakIterator& end)
    {
        std::vector<Region<DIM> > parts = split(begin, end);
        for (std::size_t s = 0; s < slabs.size(); ++s) {
            if (parts[s].size() == 0) {
                continue;
            }
            slabs[s]->setDeferSync(true);
            slabs[s]->loadMemberStreaks(source, sourceLocation, selector, parts[s].beginStreak(), parts[s].endStreak());
            slabs[s]->setDeferSync(false);
            source += selector.sizeOfExternal() * parts[s].size();
        }
        for (std::size_t s = 0; s < slabs.size(); ++s) {
            slabs[s]->sync();
        }
        dirty = true;
    }

private:
    CoordBox<DIM> box;
    CELL edgeCell;
    b200geo_group *group;
    std::vector<B200Member> members;
    std::vector<std::unique_ptr<SlabType> > slabs;
    std::vector<int> bounds;
    int cellBytes;
    bool dirty = true;

    std::size_t owner(int z) const
    {
        if (z < bounds.front() || z >= bounds.back()) {
            throw std::out_of_range("coordinate outside the grid");
        }
        std::size_t s = 0;
        while (z >= bounds[s + 1]) {
            ++s;
        }
        return s;
    }

    std::vector<Region<DIM> > split(const Region<DIM>& region, const Coord<DIM>& offset) const
    {
        std::vector<Region<DIM> > parts(slabs.size());
        for (typename Region<DIM>::StreakIterator i = region.beginStreak(); i != region.endStreak(); ++i) {
            Streak<DIM> s = *i;
            s.origin += offset;
            s.endX += offset.x();
            parts[owner(s.origin[LAST])] << s;
        }
        return parts;
    }

    std::vector<Region<DIM> > split(const typename Region<DIM>::StreakIterator& begin, const typename Region<DIM>::StreakIterator& end) const
    {
        std::vector<Region<DIM> > parts(slabs.size());
        for (typename Region<DIM>::StreakIterator i = begin; i != end; ++i) {
            parts[owner(i->origin[LAST])] << *i;
        }
        return parts;
    }
};

/* The same for grids of BoxCell containers (short-range n-body): slabs of containers along z, each a
 * B200BoxGrid on its own GPU; per sweep a slab pulls its neighbours' boundary container planes into its ghost
 * planes (b200geo_boxgroup_*) — the ghost plane is the particle migration message, BoxCell pulls from its
 * neighbourhood (storage/boxcell.h:123-138). */
template<typename PARTICLE, int N>
class B200StripedBoxGrid : public GridBase<BoxCell<FixedArray<PARTICLE, N> >, 3>
{
public:
    typedef BoxCell<FixedArray<PARTICLE, N> > CELL;
    typedef GridBase<CELL, 3> Base;
    typedef B200BoxGrid<PARTICLE, N> SlabType;
    static const int DIM = 3;

    B200StripedBoxGrid(const CoordBox<3>& box, const std::vector<int>& devices, int /* ghost width: one container */,
                       const CELL& edgeCell = CELL()) :
        Base(box.dimensions),
        box(box),
        edgeCell(edgeCell),
        group(0)
    {
        int n = (int)devices.size();
        if (n < 1) {
            throw std::invalid_argument("B200StripedBoxGrid needs at least one device");
        }
        int extent = box.dimensions[2];
        if (extent / n < 1) {
            throw std::invalid_argument("slab thinner than the ghost zone");
        }
        std::vector<b200geo_boxgrid*> handles;
        for (int s = 0; s <= n; ++s) {
            bounds.push_back(box.origin[2] + (int)(((long)extent * s) / n));
        }
        for (int s = 0; s < n; ++s) {
            CoordBox<3> slabBox = box;
            slabBox.origin[2] = bounds[s];
            slabBox.dimensions[2] = bounds[s + 1] - bounds[s];
            slabs.push_back(std::unique_ptr<SlabType>(new SlabType(slabBox, edgeCell, devices[s], n > 1 && s > 0, n > 1 && s < n - 1)));
            handles.push_back(slabs.back()->raw());
        }
        B200Helpers::check(b200geo_boxgroup_create(handles.data(), n, &group));
    }

    virtual ~B200StripedBoxGrid()
    {
        b200geo_boxgroup_destroy(group);
    }

    static int defaultGhostWidth()
    {
        return 1;
    }

    std::size_t numSlabs() const
    {
        return slabs.size();
    }

    virtual void resize(const CoordBox<3>&)
    {
        throw std::logic_error("B200StripedBoxGrid cannot be resized");
    }

    virtual void set(const Coord<3>& coord, const CELL& cell)
    {
        slabs[owner(coord.z())]->set(coord, cell);
    }

    virtual void set(const Streak<3>& streak, const CELL *cells)
    {
        slabs[owner(streak.origin.z())]->set(streak, cells);
    }

    virtual CELL get(const Coord<3>& coord) const
    {
        return slabs[owner(coord.z())]->get(coord);
    }

    virtual void get(const Streak<3>& streak, CELL *cells) const
    {
        slabs[owner(streak.origin.z())]->get(streak, cells);
    }

    virtual void setEdge(const CELL& cell)
    {
        for (std::size_t s = 0; s < slabs.size(); ++s) {
            slabs[s]->setEdge(cell);
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

    void update(unsigned firstNanoStep, unsigned sweeps)
    {
        b200geo_nbody_params p = SlabType::parameters();
        B200Helpers::check(b200geo_boxgroup_step(group, &p, firstNanoStep, sweeps));
    }

    void sync() const
    {
        B200Helpers::check(b200geo_boxgroup_sync(group));
    }

    std::pair<unsigned long long, unsigned long long> exchangeStatistics() const
    {
        uint64_t out[2] = {0, 0};
        B200Helpers::check(b200geo_boxgroup_stats(group, out));
        return std::make_pair((unsigned long long)out[0], (unsigned long long)out[1]);
    }

protected:
    virtual void saveMemberImplementation(char *, MemoryLocation::Location, const Selector<CELL>&,
                                          const typename Region<3>::StreakIterator&,
                                          const typename Region<3>::StreakIterator&) const
    {
        throw std::logic_error("B200StripedBoxGrid: containers have no selectable members");
    }

    virtual void loadMemberImplementation(const char *, MemoryLocation::Location, const Selector<CELL>&,
                                          const typename Region<3>::StreakIterator&,
                                          const typename Region<3>::StreakIterator&)
    {
        throw std::logic_error("B200StripedBoxGrid: containers have no selectable members");
    }

private:
    CoordBox<3> box;
    CELL edgeCell;
    b200geo_boxgroup *group;
    std::vector<std::unique_ptr<SlabType> > slabs;
    std::vector<int> bounds;

    std::size_t owner(int z) const
    {
        if (z < bounds.front() || z >= bounds.back()) {
            throw std::out_of_range("coordinate outside the grid");
        }
        std::size_t s = 0;
        while (z >= bounds[s + 1]) {
            ++s;
        }
        return s;
    }
};

/* which striped grid backs a cell type */
template<typename CELL>
struct B200StripedGridSelector {
    typedef B200StripedGrid<CELL> Type;
};

template<typename PARTICLE, int N>
struct B200StripedGridSelector<BoxCell<FixedArray<PARTICLE, N> > > {
    typedef B200StripedBoxGrid<PARTICLE, N> Type;
};

/* ContainerCell grids (ID-keyed cargo) live on one device: B200Simulator<ContainerCell<...> > */
template<typename CARGO, std::size_t SIZE>
struct B200StripedGridSelector<ContainerCell<CARGO, SIZE, int> > {
    static_assert(sizeof(CARGO) == 0, "there is no slab partition for ContainerCell grids yet: use B200Simulator<ContainerCell<...> >");
};

template<typename CELL>
class B200StripingSimulator : public MonolithicSimulator<CELL>
{
public:
    typedef typename MonolithicSimulator<CELL>::Topology Topology;
    typedef typename Steerer<CELL>::SteererFeedback SteererFeedback;
    typedef typename B200StripedGridSelector<CELL>::Type GridType;
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

    /* every CUDA device of the box, one slab each */
    static std::vector<int> allDevices()
    {
        int n = b200geo_device_count();
        B200Helpers::check(n);
        if (n < 1) {
            throw std::runtime_error("CUDA error: no device (the b200geo hot path has no CPU fallback)");
        }
        std::vector<int> ret;
        for (int i = 0; i < n; ++i) {
            ret.push_back(i);
        }
        return ret;
    }

    static int defaultGhostWidth()
    {
        return GridType::defaultGhostWidth();
    }

    explicit B200StripingSimulator(
        Initializer<CELL> *init,
        const std::vector<int>& devices = allDevices(),
        int ghostWidth = defaultGhostWidth()) :
        MonolithicSimulator<CELL>(init),
        grid(CoordBox<DIM>(Coord<DIM>(), init->gridBox().dimensions), devices, ghostWidth)
    {
        stepNum = init->startStep();
        simArea << CoordBox<DIM>(Coord<DIM>(), init->gridBox().dimensions);
        initializer->grid(&grid);
    }

    /* The reference's parallel simulators are constructed with a LoadBalancer:
     * StripingSimulator(initializer, balancer, loadBalancingPeriod), parallelization/stripingsimulator.h:58-75, and
     * HiParSimulator(initializer, balancer, loadBalancingPeriod, ghostZoneWidth), parallelization/hiparsimulator.h:60-66.
     * These overloads keep such call sites compiling. The slabs here are equal and the GPUs of a box identical, so the
     * balancer has nothing to decide: it is owned (deleted with the simulator, as the reference does) and never asked. */
    B200StripingSimulator(
        Initializer<CELL> *init,
        LoadBalancer *balancer,
        unsigned /* loadBalancingPeriod */ = 1,
        unsigned ghostZoneWidth = 0) :
        MonolithicSimulator<CELL>(init),
        grid(CoordBox<DIM>(Coord<DIM>(), init->gridBox().dimensions), allDevices(),
             ghostZoneWidth > 0 ? (int)ghostZoneWidth : defaultGhostWidth()),
        balancer(balancer)
    {
        stepNum = init->startStep();
        simArea << CoordBox<DIM>(Coord<DIM>(), init->gridBox().dimensions);
        initializer->grid(&grid);
    }

    using MonolithicSimulator<CELL>::addWriter;

    /* Programs written for the reference's StripingSimulator / HiParSimulator (DistributedSimulator,
     * parallelization/distributedsimulator.h:43-46) register ParallelWriters. This simulator holds the whole
     * simulation space in one process, so each of them is called once per event with the whole area as validRegion,
     * rank 0 and lastCall = true — what StripingSimulator::handleOutput does on a single rank
     * (parallelization/stripingsimulator.h:337-352). The simulator takes ownership, like the reference. Writers that
     * are BOTH a Writer and a ParallelWriter (TracingWriter, MockWriter) keep going through addWriter(Writer *). */
    template<typename WRITER>
    typename std::enable_if<std::is_base_of<ParallelWriter<CELL>, WRITER>::value &&
                            !std::is_base_of<Writer<CELL>, WRITER>::value>::type
    addWriter(WRITER *writer)
    {
        parallelWriters.push_back(typename SharedPtr<ParallelWriter<CELL> >::Type(writer));
    }

    virtual void step()
    {
        SteererFeedback feedback;
        step(&feedback, false);
    }

    virtual void run()
    {
        initializer->grid(&grid);
        stepNum = initializer->startStep();
        for (unsigned i = 0; i < steerers.size(); i++) {
            steerers[i]->setRegion(simArea);
        }
        for (std::size_t i = 0; i < parallelWriters.size(); ++i) {
            parallelWriters[i]->setRegion(simArea);
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

    const GridType& stripedGrid() const
    {
        return grid;
    }

    /* run() fuses the steps between two plugin events into one engine call, like B200Simulator */
    bool fuseSteps = true;

protected:
    GridType grid;
    Region<DIM> simArea;
    std::vector<typename SharedPtr<ParallelWriter<CELL> >::Type> parallelWriters;
    typename SharedPtr<LoadBalancer>::Type balancer;

    void step(SteererFeedback *feedback, bool fuse)
    {
        TimeTotal t(&chronometer);
        handleInput(STEERER_NEXT_STEP, feedback);

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
                grid.sync();
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
                grid.sync();
                steerers[i]->nextStep(&grid, simArea, gridDim, getStep(), event, 0, true, feedback);
            }
        }
    }
};

}

#endif
