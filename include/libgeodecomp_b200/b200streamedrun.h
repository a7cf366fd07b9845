/* The streamed run of the C++ façade: Initializer -> sweeps -> ParallelWriters pipelined along the last axis.
 *
 * A plain run() uploads the whole grid, sweeps, downloads the whole grid, one after the other: end to end it is
 * bound by two trips over the host link with the device idle. The reference's parallel simulators hand their
 * plugins SUB-BOXES of the simulation space — Initializer::grid(GridBase*) fills whatever bounding box the grid it is
 * given has (io/initializer.h:38-71), ParallelWriter::stepFinished takes a validRegion and a lastCall flag because "the
 * simulator needs to call the writer multiple times for different parts of the grid" (io/parallelwriter.h:83-99) — so
 * a simulator may cut the last axis into chunks and run
 *
 *     upload chunk c   |   level l: fused sweeps over planes [c * B - off_l, (c + 1) * B - off_l)   |   download the
 *     (Initializer)    |   (a time-skewed wavefront, b200geo_update_box_n)                           |   finished planes
 *
 * on three streams at the same time. Same kernels, same order of arithmetic per cell: results are bit-identical to
 * the plain run. The Python mirror has had this schedule since round 1 (libgeodecomp_b200/striping.py: _run_streamed,
 * where the buffer / plane invariants are spelled out); this is the same schedule in the reference's host language.
 *
 * Included by b200simulator.h (B200Grid is declared there); not a header to include on its own. */
#ifndef LIBGEODECOMP_B200_B200STREAMEDRUN_H
#define LIBGEODECOMP_B200_B200STREAMEDRUN_H

#include <libgeodecomp/io/parallelwriter.h>
#include <libgeodecomp/misc/sharedptr.h>

#include <cstring>

namespace LibGeoDecomp {

/* A box of a B200Grid as a grid of its own: what a plugin of a streamed run is handed. Selector I/O (loadMember /
 * saveMember, storage/gridbase.h:217-261) is enqueued on the window's stream and not waited for; everything else —
 * cell-by-cell access, region streams — works as well, after waiting for the device (the slow way). */
template<typename CELL>
class B200GridWindow : public GridBase<CELL, APITraits::SelectTopology<CELL>::Value::DIM>
{
public:
    typedef typename APITraits::SelectTopology<CELL>::Value Topology;
    static const int DIM = Topology::DIM;
    typedef GridBase<CELL, DIM> Base;

    B200GridWindow(B200Grid<CELL> *grid, const CoordBox<DIM>& box, void *stream) :
        Base(grid->topologicalDimensions()),
        grid(grid),
        box(box),
        stream(stream),
        wrote(false)
    {
        grid->invalidateCache();
    }

    /* cell-by-cell writes were combined on the host: ship them and wait, so that work enqueued afterwards sees them */
    void finish()
    {
        if (wrote) {
            grid->sync();
            wrote = false;
        }
    }

    virtual void resize(const CoordBox<DIM>&)
    {
        throw std::logic_error("B200GridWindow cannot be resized");
    }

    virtual void set(const Coord<DIM>& coord, const CELL& cell)
    {
        inside(coord);
        grid->set(coord, cell);
        wrote = true;
    }

    virtual void set(const Streak<DIM>& streak, const CELL *cells)
    {
        inside(streak.origin);
        grid->set(streak, cells);
        wrote = true;
    }

    virtual CELL get(const Coord<DIM>& coord) const
    {
        settle();
        return grid->get(coord);
    }

    virtual void get(const Streak<DIM>& streak, CELL *cells) const
    {
        settle();
        grid->get(streak, cells);
    }

    virtual void setEdge(const CELL& cell)
    {
        /* Initializers set the edge cell on every call: only a NEW edge cell touches the device (after waiting for
         * everything that may still read the old one) */
        if (std::memcmp(&cell, &grid->getEdge(), sizeof(CELL)) != 0) {
            grid->sync();
            B200Helpers::check(b200geo_grid_sync(grid->raw(), stream));
            grid->setEdge(cell);
        }
    }

    virtual const CELL& getEdge() const
    {
        return grid->getEdge();
    }

    virtual CoordBox<DIM> boundingBox() const
    {
        return box;
    }

    virtual void saveRegion(std::vector<char> *buffer, const Region<DIM>& region, const Coord<DIM>& offset = Coord<DIM>()) const
    {
        settle();
        grid->saveRegion(buffer, region, offset);
    }

    virtual void loadRegion(const std::vector<char>& buffer, const Region<DIM>& region, const Coord<DIM>& offset = Coord<DIM>())
    {
        settle();
        grid->loadRegion(buffer, region, offset);
    }

protected:
    virtual void saveMemberImplementation(
        char *target,
        MemoryLocation::Location targetLocation,
        const Selector<CELL>& selector,
        const typename Region<DIM>::StreakIterator& begin,
        const typename Region<DIM>::StreakIterator& end) const
    {
        IoGuard guard(grid, stream);
        grid->saveMemberStreaks(target, targetLocation, selector, begin, end);
    }

    virtual void loadMemberImplementation(
        const char *source,
        MemoryLocation::Location sourceLocation,
        const Selector<CELL>& selector,
        const typename Region<DIM>::StreakIterator& begin,
        const typename Region<DIM>::StreakIterator& end)
    {
        IoGuard guard(grid, stream);
        grid->loadMemberStreaks(source, sourceLocation, selector, begin, end);
    }

private:
    struct IoGuard {
        IoGuard(const B200Grid<CELL> *grid, void *stream) : grid(grid) { grid->setMemberIo(stream, false); }
        ~IoGuard() { grid->setMemberIo(0, true); }
        const B200Grid<CELL> *grid;
    };

    B200Grid<CELL> *grid;
    CoordBox<DIM> box;
    void *stream;
    bool wrote;

    void inside(const Coord<DIM>& c) const
    {
        if (!box.inBounds(c)) {
            throw std::out_of_range("B200GridWindow: coordinate outside the window");
        }
    }

    /* the slow way: everything this window's stream was asked to wait for has happened before the host goes on */
    void settle() const
    {
        B200Helpers::check(b200geo_grid_sync(const_cast<B200Grid<CELL>*>(grid)->raw(), stream));
        grid->invalidateCache();
    }
};

namespace B200Helpers {

template<typename BINDING>
inline auto bindsFusedBoxes(int) -> decltype(&BINDING::updateBoxN, true)
{
    return true;
}

template<typename BINDING>
inline bool bindsFusedBoxes(long)
{
    return false;
}

}

/* The schedule. plan() says whether a run from `first` to `last` (steps) can be streamed at all and how. */
template<typename CELL>
class B200StreamedRun
{
public:
    typedef typename APITraits::SelectTopology<CELL>::Value Topology;
    static const int DIM = Topology::DIM;
    static const int LAST = DIM - 1;
    static const unsigned NANO_STEPS = APITraits::SelectNanoSteps<CELL>::VALUE;
    typedef std::vector<typename SharedPtr<ParallelWriter<CELL> >::Type> WriterVector;

    std::vector<unsigned> levels;   /* fused sweeps per level; the shorter remainder level goes first */
    int chunk;                      /* planes (rows in 2-D) per chunk */

    B200StreamedRun() : chunk(0) {}

    /* chunks = how many pieces the last axis is cut into (more pieces: shorter pipeline fill and drain, more launches) */
    bool plan(const CoordBox<DIM>& box, unsigned first, unsigned last, const WriterVector& writers, int chunks = 16)
    {
        levels.clear();
        chunk = 0;
        if (last <= first || Topology::wrapsAxis(LAST) || !B200Helpers::bindsFusedBoxes<B200KernelBinding<CELL> >(0)) {
            return false;
        }
        for (std::size_t i = 0; i < writers.size(); ++i) {
            /* no WRITER_STEP_FINISHED may fall due inside the run: no step s in (first, last) with s % period == 0 */
            unsigned p = writers[i]->getPeriod();
            if (first / p != (last - 1) / p) {
                return false;
            }
        }
        unsigned sweeps = (last - first) * NANO_STEPS;
        unsigned depth = (std::min)(B200Helpers::fusedSweeps(B200KernelBinding<CELL>::kernel()), sweeps);
        if (sweeps % depth) {
            levels.push_back(sweeps % depth);
        }
        levels.insert(levels.end(), sweeps / depth, depth);
        int n = box.dimensions[LAST];
        int c = (std::max)(2 * (int)depth, (n + chunks - 1) / (std::max)(1, chunks));
        c = (c + (int)depth - 1) / (int)depth * (int)depth;   /* boxes cut off at plane 0 stay whole multiples of the depth */
        if (n < 2 * c) {
            levels.clear();
            return false;
        }
        chunk = c;
        return true;
    }

    /* Runs the schedule on `grid`. Returns the number of box launches. */
    template<typename INITIALIZER>
    std::size_t run(B200Grid<CELL> *grid, INITIALIZER *initializer, const WriterVector& writers, const Coord<DIM>& globalDimensions) const
    {
        const CoordBox<DIM> box = grid->boundingBox();
        const int n = box.dimensions[LAST];
        const int L = (int)levels.size();
        std::vector<int> off(L);
        std::vector<unsigned> firstNano(L);
        int sum = 0;
        for (int l = 0; l < L; ++l) {
            firstNano[l] = (unsigned)sum % NANO_STEPS;
            sum += (int)levels[l];
            off[l] = sum;
        }
        const unsigned first = initializer->startStep(), last = initializer->maxSteps();
        b200geo_grid *handle = grid->raw();
        const int device = grid->deviceIndex();
        Streams streams(device);
        int parity = 0;
        std::size_t launches = 0;
        const int uploads = (n + chunk - 1) / chunk;
        const int rounds = (n + off[L - 1] + chunk - 1) / chunk;
        for (int c = 0; c < rounds; ++c) {
            if (c < uploads) {
                int a = c * chunk, b = (std::min)((c + 1) * chunk, n);
                want(handle, &parity, 0);
                B200GridWindow<CELL> window(grid, planes(box, a, b), streams.up);
                initializer->grid(&window);
                window.finish();
                Region<DIM> region;
                region << planes(box, a, b);
                for (std::size_t i = 0; i < writers.size(); ++i) {
                    writers[i]->stepFinished(window, region, globalDimensions, first, WRITER_INITIALIZED, 0, b == n);
                }
                B200Helpers::check(b200geo_stream_wait(device, streams.run, streams.up));
            }
            for (int l = 0; l < L; ++l) {
                int a = (std::max)(c * chunk - off[l], 0), b = (std::min)((c + 1) * chunk - off[l], n);
                if (b <= a) {
                    continue;
                }
                want(handle, &parity, l % 2);
                int32_t origin[3] = {0, 0, 0}, dim[3] = {1, 1, 1};
                for (int i = 0; i < DIM; ++i) {
                    dim[i] = box.dimensions[i];
                }
                origin[LAST] = a;
                dim[LAST] = b - a;
                launch<B200KernelBinding<CELL> >(0, handle, firstNano[l], origin, dim, levels[l], l == L - 1, streams.run);
                ++launches;
            }
            int a = (std::max)(c * chunk - off[L - 1], 0), b = (std::min)((c + 1) * chunk - off[L - 1], n);
            if (b > a && !writers.empty()) {
                B200Helpers::check(b200geo_stream_wait(device, streams.down, streams.run));
                want(handle, &parity, L % 2);
                B200GridWindow<CELL> window(grid, planes(box, a, b), streams.down);
                Region<DIM> region;
                region << planes(box, a, b);
                for (std::size_t i = 0; i < writers.size(); ++i) {
                    writers[i]->stepFinished(window, region, globalDimensions, last, WRITER_ALL_DONE, 0, b == n);
                }
            }
        }
        want(handle, &parity, L % 2);   /* the final state is the current buffer from here on */
        streams.join(handle);
        grid->invalidateCache();
        return launches;
    }

private:
    /* cells whose binding has no box updates never get here (plan() says no); the call must still compile for them */
    template<typename BINDING>
    static auto launch(int, b200geo_grid *handle, unsigned nanoStep, const int32_t *origin, const int32_t *dim, unsigned sweeps,
                       bool final, void *stream) -> decltype(BINDING::updateBoxN(handle, nanoStep, origin, dim, sweeps, final, stream), void())
    {
        BINDING::updateBoxN(handle, nanoStep, origin, dim, sweeps, final, stream);
    }

    template<typename BINDING>
    static void launch(long, b200geo_grid *, unsigned, const int32_t *, const int32_t *, unsigned, bool, void *)
    {
        throw std::logic_error("B200StreamedRun: this cell's binding does not update boxes");
    }

    struct Streams {
        explicit Streams(int device) : device(device), up(0), run(0), down(0)
        {
            B200Helpers::check(b200geo_stream_create(device, &up));
            B200Helpers::check(b200geo_stream_create(device, &run));
            B200Helpers::check(b200geo_stream_create(device, &down));
        }

        ~Streams()
        {
            b200geo_stream_destroy(device, up);
            b200geo_stream_destroy(device, run);
            b200geo_stream_destroy(device, down);
        }

        /* the host waits for all three: the ParallelWriters' buffers are complete when run() returns */
        void join(b200geo_grid *handle)
        {
            B200Helpers::check(b200geo_grid_sync(handle, up));
            B200Helpers::check(b200geo_grid_sync(handle, run));
            B200Helpers::check(b200geo_grid_sync(handle, down));
        }

        int device;
        void *up, *run, *down;
    };

    /* which buffer the C ABI calls "current" is a host-side flag; device work already enqueued keeps its pointers */
    static void want(b200geo_grid *handle, int *parity, int p)
    {
        if (*parity != p) {
            B200Helpers::check(b200geo_swap(handle));
            *parity = p;
        }
    }

    static CoordBox<DIM> planes(const CoordBox<DIM>& box, int a, int b)
    {
        CoordBox<DIM> ret = box;
        ret.origin[LAST] = box.origin[LAST] + a;
        ret.dimensions[LAST] = b - a;
        return ret;
    }
};

}

#endif
