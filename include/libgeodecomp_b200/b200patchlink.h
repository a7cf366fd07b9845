/* B200PatchLink<CELL>: a ghost zone link between two B200Steppers of one process that never leaves the GPUs.
 *
 * The reference ships ghost zones between the Steppers of different MPI ranks through PatchLink::Accepter (put: save
 * the region, MPI_Isend) and PatchLink::Provider (MPI_Irecv, get: load the region), communication/patchlink.h:127-151,
 * 218-244; inside one process its tests link steppers with PatchBuffer (storage/patchbuffer.h:19-80), a
 * PatchAccepter + PatchProvider that keeps the saved regions in host memory. Both see the steppers' HOST grids.
 *
 * This link is registered exactly like a PatchBuffer — as an accepter of the sending stepper and a provider of the
 * receiving one, charged with pushRequest() — but when both ends are B200Steppers (b200stepper.h recognises the
 * B200DevicePatchAccepter / B200DevicePatchProvider interfaces) the region goes
 *     sender's device grid --one region-copy launch--> buffer on the sender's GPU
 *                          --cudaMemcpyPeer (NVLink between peer GPUs; nothing if both steppers share a GPU)-->
 *                          --one region-copy launch--> receiver's device grid
 * and the host grids are not touched. The nano step bookkeeping is PatchBuffer's (requested / stored nano steps), so
 * the steppers' schedule is unchanged. Ends that are not B200Steppers still work: put(hostGrid) / get(hostGrid) keep
 * the region in host memory, as PatchBuffer does, and a B200Stepper at the other end takes it from / hands it to there.
 *
 * All device work is enqueued on the null streams of the two devices (as the steppers' own launches are), and
 * b200geo_device_copy is ordered after the sender's launches and before the receiver's: no explicit synchronisation. */
#ifndef LIBGEODECOMP_B200_B200PATCHLINK_H
#define LIBGEODECOMP_B200_B200PATCHLINK_H

#include <libgeodecomp/parallelization/nesting/stepper.h>
#include <libgeodecomp/storage/patchaccepter.h>
#include <libgeodecomp/storage/patchprovider.h>
#include <libgeodecomp/storage/serializationbuffer.h>

#include <deque>

#include "b200simulator.h"

namespace LibGeoDecomp {

/* what B200Stepper looks for among its PatchAccepters / PatchProviders: ends that take the DEVICE grid */
template<typename CELL>
class B200DevicePatchAccepter
{
public:
    static const int DIM = APITraits::SelectTopology<CELL>::Value::DIM;

    virtual ~B200DevicePatchAccepter() {}

    virtual void putDevice(const B200Grid<CELL>& grid, const Region<DIM>& validRegion, const Coord<DIM>& globalGridDimensions,
                           std::size_t nanoStep, std::size_t rank) = 0;
};

template<typename CELL>
class B200DevicePatchProvider
{
public:
    static const int DIM = APITraits::SelectTopology<CELL>::Value::DIM;

    virtual ~B200DevicePatchProvider() {}

    virtual void getDevice(B200Grid<CELL> *grid, const Region<DIM>& patchableRegion, const Coord<DIM>& globalGridDimensions,
                           std::size_t nanoStep, std::size_t rank, bool remove) = 0;
};

/* GRID = the host grid type of the steppers (Stepper<CELL>::GridType, parallelization/nesting/stepper.h: what their
 * PatchAccepterVec / PatchProviderVec are declared for) */
template<typename CELL, typename GRID = typename Stepper<CELL>::GridType>
class B200PatchLink :
        public PatchAccepter<GRID>,
        public PatchProvider<GRID>,
        public B200DevicePatchAccepter<CELL>,
        public B200DevicePatchProvider<CELL>
{
public:
    static const int DIM = APITraits::SelectTopology<CELL>::Value::DIM;
    typedef GRID GridType;
    typedef typename SerializationBuffer<CELL>::BufferType BufferType;

    using PatchAccepter<GridType>::checkNanoStepPut;
    using PatchAccepter<GridType>::requestedNanoSteps;
    using PatchProvider<GridType>::checkNanoStepGet;
    using PatchProvider<GridType>::storedNanoSteps;

    explicit B200PatchLink(const Region<DIM>& region = Region<DIM>()) :
        region(region),
        deviceTransfers(0),
        peerCopies(0),
        hostTransfers(0)
    {}

    virtual ~B200PatchLink()
    {
        for (typename std::deque<Stored>::iterator i = stored.begin(); i != stored.end(); ++i) {
            release(*i);
        }
        for (std::size_t i = 0; i < pool.size(); ++i) {
            b200geo_device_free(pool[i].device, pool[i].data);
        }
    }

    /* ---- the sending end */
    virtual void putDevice(const B200Grid<CELL>& grid, const Region<DIM>&, const Coord<DIM>&, std::size_t nanoStep, std::size_t)
    {
        if (!checkNanoStepPut(nanoStep)) {
            return;
        }
        Stored s;
        s.device = grid.deviceIndex();
        s.bytes = region.size() * (std::size_t)grid.bytesPerCell();
        s.data = acquire(s.device, s.bytes);
        grid.saveRegionToDevice(s.data, region);
        stored.push_back(s);
        storedNanoSteps << (min)(requestedNanoSteps);
        erase_min(requestedNanoSteps);
        ++deviceTransfers;
    }

    virtual void put(const GridType& grid, const Region<DIM>&, const Coord<DIM>&, const std::size_t nanoStep, const std::size_t)
    {
        if (!checkNanoStepPut(nanoStep)) {
            return;
        }
        Stored s;
        s.device = -1;
        s.data = 0;
        s.bytes = 0;
        s.host = SerializationBuffer<CELL>::create(region);
        grid.saveRegion(&s.host, region);
        stored.push_back(s);
        storedNanoSteps << (min)(requestedNanoSteps);
        erase_min(requestedNanoSteps);
        ++hostTransfers;
    }

    /* ---- the receiving end */
    virtual void getDevice(B200Grid<CELL> *grid, const Region<DIM>&, const Coord<DIM>&, std::size_t nanoStep, std::size_t, bool remove)
    {
        checkNanoStepGet(nanoStep);
        if (stored.empty()) {
            throw std::logic_error("no region available");
        }
        Stored& s = stored.front();
        if (s.device < 0) {
            loadFromHost(grid, s.host);
        } else if (s.device == grid->deviceIndex()) {
            grid->loadRegionFromDevice(s.data, region);
        } else {
            /* another GPU of the box: one peer copy into a buffer on the receiver's GPU */
            void *here = acquire(grid->deviceIndex(), s.bytes);
            B200Helpers::check(b200geo_device_copy(grid->deviceIndex(), here, s.device, s.data, s.bytes));
            grid->loadRegionFromDevice(here, region);
            Buffer back = {here, grid->deviceIndex(), s.bytes};
            pool.push_back(back);
            ++peerCopies;
        }
        if (remove) {
            release(s);
            stored.pop_front();
            erase_min(storedNanoSteps);
        }
    }

    virtual void get(GridType *destinationGrid, const Region<DIM>&, const Coord<DIM>&, const std::size_t nanoStep, const std::size_t,
                     const bool remove = true)
    {
        checkNanoStepGet(nanoStep);
        if (stored.empty()) {
            throw std::logic_error("no region available");
        }
        Stored& s = stored.front();
        if (s.device < 0) {
            destinationGrid->loadRegion(s.host, region);
        } else {
            BufferType buffer = SerializationBuffer<CELL>::create(region);
            readBack(&buffer, s);
            destinationGrid->loadRegion(buffer, region);
        }
        if (remove) {
            release(s);
            stored.pop_front();
            erase_min(storedNanoSteps);
        }
    }

    /* regions that went device to device / of those, across two GPUs / through host memory */
    std::size_t deviceTransferCount() const { return deviceTransfers; }
    std::size_t peerCopyCount() const { return peerCopies; }
    std::size_t hostTransferCount() const { return hostTransfers; }

private:
    struct Stored {
        void *data;          /* device buffer (member-major stream of the region), or ... */
        int device;          /* ... -1: */
        std::size_t bytes;
        BufferType host;     /* the region as the host grid serialised it */
    };

    struct Buffer {
        void *data;
        int device;
        std::size_t bytes;
    };

    Region<DIM> region;
    std::deque<Stored> stored;
    std::vector<Buffer> pool;    /* device buffers are kept for the next nano step: no cudaMalloc per transfer */
    std::size_t deviceTransfers, peerCopies, hostTransfers;

    void *acquire(int device, std::size_t bytes)
    {
        for (std::size_t i = 0; i < pool.size(); ++i) {
            if (pool[i].device == device && pool[i].bytes >= bytes) {
                void *ret = pool[i].data;
                pool.erase(pool.begin() + i);
                return ret;
            }
        }
        void *ret = 0;
        B200Helpers::check(b200geo_device_alloc(device, bytes, &ret));
        return ret;
    }

    void release(Stored& s)
    {
        if (s.device >= 0 && s.data) {
            Buffer b = {s.data, s.device, s.bytes};
            pool.push_back(b);
            s.data = 0;
        }
    }

    static void loadFromHost(B200Grid<CELL> *grid, const std::vector<char>& buffer, const Region<DIM>& region)
    {
        grid->loadRegionBytes(buffer, region, Coord<DIM>(), 0);
    }

    static void loadFromHost(B200Grid<CELL> *grid, const std::vector<CELL>& buffer, const Region<DIM>& region)
    {
        grid->loadRegionCells(buffer, region, Coord<DIM>(), 0);
    }

    void loadFromHost(B200Grid<CELL> *grid, const BufferType& buffer)
    {
        loadFromHost(grid, buffer, region);
    }

    /* a region a B200Stepper left on its GPU, asked for by a stepper with a host grid: SoA cells travel as the same
     * member-major bytes on both sides */
    void readBack(std::vector<char> *buffer, const Stored& s)
    {
        buffer->resize(s.bytes);
        B200Helpers::check(b200geo_device_copy(-1, buffer->data(), s.device, s.data, s.bytes));
    }

    void readBack(std::vector<CELL> *, const Stored&)
    {
        throw std::logic_error("B200PatchLink: a region stored on the device can only be handed to a host grid for cells "
                               "that serialise member by member (SoA cells); link AoS cells to B200Steppers on both ends");
    }
};

}

#endif
