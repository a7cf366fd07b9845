/* B200Stepper (include/libgeodecomp_b200/b200stepper.h) beside the reference's VanillaStepper, set up the way the
 * reference tests its steppers (parallelization/nesting/test/parallel_mpi_1/vanillastepperregiontest.h:48-88): a
 * hand-built PartitionManager that makes this process "rank 1 of 3" of a StripingPartition with ragged weights
 * (region borders in the middle of rows), no MPI anywhere. Both steppers get
 *   - a ghost zone PatchProvider that hands them the neighbours' cells of the outer ghost zone at every
 *     synchronisation point — taken from a SerialSimulator run of the WHOLE simulation space, i.e. what real
 *     neighbours would send;
 *   - a ghost zone PatchAccepter and an inner set PatchAccepter that record what the stepper hands out.
 * After every update1() the stepper's grid must equal the whole-space run on the inner set that is valid at that
 * nano step (vanillastepperregiontest.h:114-124 checks exactly those sets), the records of the two steppers must be
 * identical, and B200Stepper must have moved regions between host and device only when a patch was due.
 * Ghost zone widths 1-4, 2-D (Game of Life) and 3-D (Jacobi 7- and 27-point).
 * Linked against libb200geo.so (GPU box) or tests/facade/mock_b200geo.cpp (CPU suite). */
#include <libgeodecomp/geometry/partitionmanager.h>
#include <libgeodecomp/geometry/partitions/recursivebisectionpartition.h>
#include <libgeodecomp/geometry/partitions/stripingpartition.h>
#include <libgeodecomp/storage/patchbuffer.h>
#include <libgeodecomp/parallelization/nesting/vanillastepper.h>
#include <libgeodecomp/storage/patchaccepter.h>
#include <libgeodecomp/storage/patchprovider.h>

#include <libgeodecomp_b200/b200stepper.h>

#include "fixtures.h"

/* one patch a stepper handed to an accepter */
template<typename CELL>
struct Record {
    std::size_t nanoStep;
    std::size_t cells;
    std::vector<CELL> data;
};

template<typename GRID>
class RecordingAccepter : public PatchAccepter<GRID>
{
public:
    typedef typename GRID::CellType CELL;
    static const int DIM = GRID::DIM;
    using PatchAccepter<GRID>::requestedNanoSteps;
    using PatchAccepter<GRID>::pushRequest;

    RecordingAccepter(std::size_t first, std::size_t stride, std::size_t last)
    {
        for (std::size_t s = first; s <= last; s += stride) {
            pushRequest(s);
        }
    }

    virtual void put(const GRID& grid, const Region<DIM>& validRegion, const Coord<DIM>&, const std::size_t nanoStep, const std::size_t)
    {
        if (!this->checkNanoStepPut(nanoStep)) {
            return;
        }
        Record<CELL> r;
        r.nanoStep = nanoStep;
        r.cells = validRegion.size();
        for (typename Region<DIM>::Iterator i = validRegion.begin(); i != validRegion.end(); ++i) {
            r.data.push_back(grid.get(*i));
        }
        records.push_back(r);
        requestedNanoSteps.erase(requestedNanoSteps.begin());
    }

    std::vector<Record<CELL> > records;
};

/* the neighbours: cells of the outer ghost zone out of the whole-space run */
template<typename GRID, typename CELL>
class NeighbourProvider : public PatchProvider<GRID>
{
public:
    static const int DIM = GRID::DIM;
    using PatchProvider<GRID>::storedNanoSteps;

    NeighbourProvider(const std::vector<std::vector<CELL> > *history, const CoordBox<DIM>& box, const Region<DIM>& outerGhost,
                      std::size_t first, std::size_t stride, std::size_t last) :
        history(history), box(box), outerGhost(outerGhost)
    {
        for (std::size_t s = first; s <= last; s += stride) {
            storedNanoSteps.insert(s);
        }
    }

    virtual void get(GRID *grid, const Region<DIM>& patchableRegion, const Coord<DIM>&, const std::size_t nanoStep, const std::size_t, const bool remove = true)
    {
        this->checkNanoStepGet(nanoStep);
        Region<DIM> region = outerGhost & patchableRegion;
        for (typename Region<DIM>::Iterator i = region.begin(); i != region.end(); ++i) {
            grid->set(*i, (*history)[nanoStep][(*i - box.origin).toIndex(box.dimensions)]);
        }
        if (remove) {
            storedNanoSteps.erase(storedNanoSteps.begin());
        }
    }

private:
    const std::vector<std::vector<CELL> > *history;
    CoordBox<DIM> box;
    Region<DIM> outerGhost;
};

/* The reference test gives rank 1 two rows (vanillastepperregiontest.h:57-61: 4 rows + 7 cells | 2 rows - 1 cell | the
 * rest) — with ghost zone width 3 its kernel and all its inner sets are EMPTY, so that test checks no cell. Here rank 1
 * gets 2 * ghostZoneWidth + 4 rows / planes (minus a cell), region borders still in the middle of rows. */
template<int DIM> struct Weights;
template<> struct Weights<2> {
    static std::vector<std::size_t> make(const Coord<2>& dim, unsigned width)
    {
        std::vector<std::size_t> w(3);
        w[0] = 4 * dim.x() + 7;
        w[1] = (2 * width + 4) * dim.x() - 1;
        w[2] = dim.prod() - w[0] - w[1];
        return w;
    }
};
template<> struct Weights<3> {
    static std::vector<std::size_t> make(const Coord<3>& dim, unsigned width)
    {
        std::vector<std::size_t> w(3);
        std::size_t plane = (std::size_t)dim.x() * dim.y();
        w[0] = 3 * plane + 2 * dim.x() + 5;       /* ends in the middle of a row */
        w[1] = (2 * width + 4) * plane + 3 * dim.x() - 2;
        w[2] = dim.prod() - w[0] - w[1];
        return w;
    }
};

template<typename CELL>
static void runCase(const char *name, const Coord<APITraits::SelectTopology<CELL>::Value::DIM>& dim, unsigned ghostZoneWidth, unsigned steps)
{
    typedef typename APITraits::SelectTopology<CELL>::Value Topology;
    const int DIM = Topology::DIM;
    typedef VanillaStepper<CELL, UpdateFunctorHelpers::ConcurrencyNoP> Reference;
    typedef B200Stepper<CELL> Device;
    typedef typename Reference::GridType GridType;

    /* the whole-space run: history[t] = every cell at nano step t */
    CoordBox<DIM> box(Coord<DIM>(), dim);
    std::vector<std::vector<CELL> > history;
    {
        SerialSimulator<CELL> whole(new SeededInitializer<CELL>(dim, steps + 2 * ghostZoneWidth));
        for (unsigned t = 0; t <= steps + 2 * ghostZoneWidth; ++t) {
            std::vector<CELL> now;
            for (typename CoordBox<DIM>::Iterator i = box.begin(); i != box.end(); ++i) {
                now.push_back(whole.getGrid()->get(*i));
            }
            history.push_back(now);
            whole.step();
        }
    }

    struct Setup {
        typename SharedPtr<PartitionManager<Topology> >::Type manager;
        typename SharedPtr<SeededInitializer<CELL> >::Type init;
        typename SharedPtr<RecordingAccepter<GridType> >::Type ghostAccepter, innerAccepter;
        typename SharedPtr<NeighbourProvider<GridType, CELL> >::Type provider;
    };
    Setup setups[2];
    const std::size_t last = steps + 2 * ghostZoneWidth;
    for (int s = 0; s < 2; ++s) {
        Setup& u = setups[s];
        u.init.reset(new SeededInitializer<CELL>(dim, last));
        typename SharedPtr<Partition<DIM> >::Type partition(new StripingPartition<DIM>(Coord<DIM>(), dim, 0, Weights<DIM>::make(dim, ghostZoneWidth)));
        typename SharedPtr<AdjacencyManufacturer<DIM> >::Type adjacency(new DummyAdjacencyManufacturer<DIM>);
        u.manager.reset(new PartitionManager<Topology>());
        u.manager->resetRegions(adjacency, box, partition, 1, ghostZoneWidth);
        std::vector<CoordBox<DIM> > boundingBoxes, expandedBoundingBoxes;
        for (int i = 0; i < 3; ++i) {
            Region<DIM> region = partition->getRegion(i);
            boundingBoxes.push_back(region.boundingBox());
            expandedBoundingBoxes.push_back(region.expandWithTopology(ghostZoneWidth, dim, Topology()).boundingBox());
        }
        u.manager->resetGhostZones(boundingBoxes, expandedBoundingBoxes);
        /* synchronisation points as UpdateGroup charges its PatchLinks: ghostZoneWidth, 2 * ghostZoneWidth, ... */
        u.ghostAccepter.reset(new RecordingAccepter<GridType>(ghostZoneWidth, ghostZoneWidth, last));
        u.innerAccepter.reset(new RecordingAccepter<GridType>(3, 3, last));     /* a writer with period 3 */
        u.provider.reset(new NeighbourProvider<GridType, CELL>(&history, box, u.manager->getOuterRim(), ghostZoneWidth, ghostZoneWidth, last));
    }

    typename Reference::PatchAccepterVec ghostAccepters[2], innerAccepters[2];
    typename Reference::PatchProviderVec providers[2];
    for (int s = 0; s < 2; ++s) {
        ghostAccepters[s].push_back(setups[s].ghostAccepter);
        innerAccepters[s].push_back(setups[s].innerAccepter);
        providers[s].push_back(setups[s].provider);
    }
    Reference reference(setups[0].manager, setups[0].init, ghostAccepters[0], innerAccepters[0], providers[0]);
    Device device(setups[1].manager, setups[1].init, ghostAccepters[1], innerAccepters[1], providers[1]);

    long wrongVsWhole = 0, wrongVsReference = 0, compared = 0;
    std::size_t pullsWithoutPatch = 0;
    for (unsigned t = 1; t <= steps; ++t) {
        std::size_t pullsBefore = device.pullCount(), recordsBefore = setups[1].ghostAccepter->records.size() + setups[1].innerAccepter->records.size();
        std::size_t providerBefore = setups[1].provider->nextAvailableNanoStep();
        reference.update(1);
        device.update(1);
        bool patchDue = setups[1].ghostAccepter->records.size() + setups[1].innerAccepter->records.size() != recordsBefore ||
                        setups[1].provider->nextAvailableNanoStep() != providerBefore;
        if (!patchDue && device.pullCount() != pullsBefore) {
            ++pullsWithoutPatch;
        }
        CHECK(device.currentStep() == reference.currentStep());
        /* the inner set that is valid now: shrunk by the nano steps since the last synchronisation */
        unsigned shrink = t % ghostZoneWidth;
        const Region<DIM>& valid = setups[0].manager->innerSet(shrink);
        const GridType& want = reference.grid();
        const GridType& got = device.grid();
        for (typename Region<DIM>::Iterator i = valid.begin(); i != valid.end(); ++i) {
            CELL g = got.get(*i);
            ++compared;
            if (!(g == want.get(*i))) ++wrongVsReference;
            if (!(g == history[t][(*i - box.origin).toIndex(box.dimensions)])) ++wrongVsWhole;
        }
    }
    CHECK(wrongVsWhole == 0);
    CHECK(wrongVsReference == 0);
    CHECK(compared > 0);
    CHECK(pullsWithoutPatch == 0);

    bool sameRecords = true;
    std::size_t patches = 0;
    for (int which = 0; which < 2; ++which) {
        const std::vector<Record<CELL> >& a = which ? setups[0].innerAccepter->records : setups[0].ghostAccepter->records;
        const std::vector<Record<CELL> >& b = which ? setups[1].innerAccepter->records : setups[1].ghostAccepter->records;
        sameRecords &= a.size() == b.size() && !a.empty();
        for (std::size_t k = 0; sameRecords && k < a.size(); ++k) {
            sameRecords &= a[k].nanoStep == b[k].nanoStep && a[k].cells == b[k].cells && a[k].data.size() == b[k].data.size();
            for (std::size_t c = 0; sameRecords && c < a[k].data.size(); ++c) {
                sameRecords &= a[k].data[c] == b[k].data[c];
            }
            ++patches;
        }
    }
    CHECK(sameRecords);
    std::printf("%-13s ghost zone width %u, %u nano steps as rank 1 of 3: %ld / %ld of %ld cells differ from the whole-space run / VanillaStepper, "
                "%zu patches %s, %zu launches, %zu pulls, %zu pushes\n",
                name, ghostZoneWidth, steps, wrongVsWhole, wrongVsReference, compared, patches, sameRecords ? "identical" : "DIFFERENT",
                device.launchCount(), device.pullCount(), device.pushCount());
}

/* Bricks: the simulation space cut by a RecursiveBisectionPartition (geometry/partitions/recursivebisectionpartition.h:
 * 17,103-127 — what north_star calls the 2 x 2 x 2 brick partition) into `nodes` cuboids, ONE B200Stepper per brick, the
 * ghost zones shipped between them through the reference's own in-memory PatchAccepter + PatchProvider, PatchBuffer
 * (storage/patchbuffer.h:19-80) — wired the way UpdateGroup wires its PatchLinks (parallelization/nesting/updategroup.h:
 * inner ghost zone fragment -> accepter on the sender, outer ghost zone fragment -> provider on the receiver, charged with
 * ghostZoneWidth, 2 * ghostZoneWidth, ...). The steppers are advanced round-robin, ghostZoneWidth nano steps each; the
 * union of their own regions must equal the whole-space SerialSimulator run. No MPI, no host-side simulator. */
template<typename LINK> struct LinkStats {
    static std::size_t device(const LINK&) { return 0; }
    static std::size_t peer(const LINK&) { return 0; }
    static const char *name() { return "PatchBuffer"; }
};
template<typename CELL, typename GRID> struct LinkStats<B200PatchLink<CELL, GRID> > {
    static std::size_t device(const B200PatchLink<CELL, GRID>& link) { return link.deviceTransferCount(); }
    static std::size_t peer(const B200PatchLink<CELL, GRID>& link) { return link.peerCopyCount(); }
    static const char *name() { return "B200PatchLink"; }
};

/* devices the bricks' steppers run on: every GPU of the box round-robin; the mock engine (one pretend device) is told
 * about two, so that the copy between the GPUs of a link is part of the CPU suite as well */
static int deviceFor(int node)
{
    static int count = 0;
    if (count == 0) {
        count = std::string(b200geo_version()).find("mock") != std::string::npos ? 2 : (std::max)(1, b200geo_device_count());
    }
    return node % count;
}

template<typename CELL, typename LINK>
static void runBricksWith(const char *name, const Coord<3>& dim, int nodes, unsigned ghostZoneWidth, unsigned rounds)
{
    typedef typename APITraits::SelectTopology<CELL>::Value Topology;
    typedef B200Stepper<CELL> StepperType;
    typedef typename StepperType::GridType GridType;
    typedef LINK Link;
    const unsigned steps = ghostZoneWidth * rounds;
    CoordBox<3> box(Coord<3>(), dim);

    SerialSimulator<CELL> whole(new SeededInitializer<CELL>(dim, steps));
    whole.run();

    std::vector<std::size_t> weights(nodes, dim.prod() / nodes);
    weights.back() += dim.prod() % nodes;
    typename SharedPtr<Partition<3> >::Type partition(new RecursiveBisectionPartition<3>(Coord<3>(), dim, 0, weights));
    std::vector<CoordBox<3> > boundingBoxes, expandedBoundingBoxes;
    bool cuboids = true;
    for (int i = 0; i < nodes; ++i) {
        Region<3> region = partition->getRegion(i);
        cuboids &= region.size() == (std::size_t)region.boundingBox().dimensions.prod();
        boundingBoxes.push_back(region.boundingBox());
        expandedBoundingBoxes.push_back(region.expandWithTopology(ghostZoneWidth, dim, Topology()).boundingBox());
    }
    CHECK(cuboids);

    std::vector<typename SharedPtr<PartitionManager<Topology> >::Type> managers(nodes);
    for (int i = 0; i < nodes; ++i) {
        typename SharedPtr<AdjacencyManufacturer<3> >::Type adjacency(new DummyAdjacencyManufacturer<3>);
        managers[i].reset(new PartitionManager<Topology>());
        managers[i]->resetRegions(adjacency, box, partition, i, ghostZoneWidth);
        managers[i]->resetGhostZones(boundingBoxes, expandedBoundingBoxes);
    }
    /* link (i -> j): i's cells inside j's ghost zone */
    std::vector<typename StepperType::PatchAccepterVec> accepters(nodes);
    std::vector<typename StepperType::PatchProviderVec> providers(nodes);
    std::size_t links = 0;
    std::vector<typename SharedPtr<Link>::Type> allLinks;
    for (int i = 0; i < nodes; ++i) {
        typename PartitionManager<Topology>::RegionVecMap& inner = managers[i]->getInnerGhostZoneFragments();
        for (typename PartitionManager<Topology>::RegionVecMap::iterator f = inner.begin(); f != inner.end(); ++f) {
            int j = f->first;
            if (j < 0 || j >= nodes || f->second.back().empty()) {
                continue;     /* OUTGROUP pseudo node */
            }
            const Region<3>& sent = f->second.back();
            const Region<3>& received = managers[j]->getOuterGhostZoneFragments()[i].back();
            CHECK(sent == received);
            typename SharedPtr<Link>::Type link(new Link(sent));
            for (unsigned t = ghostZoneWidth; t <= steps + ghostZoneWidth; t += ghostZoneWidth) {
                link->pushRequest(t);
            }
            accepters[i].push_back(link);
            providers[j].push_back(link);
            allLinks.push_back(link);
            ++links;
        }
    }
    std::vector<typename SharedPtr<StepperType>::Type> steppers(nodes);
    for (int i = 0; i < nodes; ++i) {
        typename SharedPtr<SeededInitializer<CELL> >::Type init(new SeededInitializer<CELL>(dim, steps));
        steppers[i].reset(new StepperType(managers[i], init, accepters[i], typename StepperType::PatchAccepterVec(), providers[i],
                                          typename StepperType::PatchProviderVec(), typename StepperType::PatchProviderVec(), false,
                                          deviceFor(i)));
    }
    for (unsigned r = 0; r < rounds; ++r) {
        for (int i = 0; i < nodes; ++i) {
            steppers[i]->update(ghostZoneWidth);
        }
    }
    long bad = 0, cells = 0;
    std::size_t launches = 0, pullsBefore = 0, direct = 0, peer = 0, puts = 0, gets = 0;
    for (int i = 0; i < nodes; ++i) {
        pullsBefore += steppers[i]->pullCount();
        puts += steppers[i]->devicePutCount();
        gets += steppers[i]->deviceGetCount();
    }
    for (std::size_t k = 0; k < allLinks.size(); ++k) {
        direct += LinkStats<Link>::device(*allLinks[k]);
        peer += LinkStats<Link>::peer(*allLinks[k]);
    }
    const bool deviceLinks = std::string(LinkStats<Link>::name()) == "B200PatchLink";
    if (deviceLinks) {
        /* every ghost zone went device to device: the host grids saw nothing but the initial state */
        CHECK(direct == links * (rounds + 1));
        CHECK(puts == direct && gets == links * rounds);
        CHECK(pullsBefore == 0);
    }
    for (int i = 0; i < nodes; ++i) {
        const GridType& grid = steppers[i]->grid();
        const Region<3>& own = managers[i]->ownRegion();
        for (typename Region<3>::Iterator c = own.begin(); c != own.end(); ++c) {
            bad += !(grid.get(*c) == whole.getGrid()->get(*c));
            ++cells;
        }
        launches += steppers[i]->launchCount();
        CHECK(steppers[i]->currentStep().first == steps);
    }
    CHECK(bad == 0);
    CHECK(cells == (long)dim.prod());
    std::printf("%-13s %d bricks (RecursiveBisectionPartition), ghost zone width %u, %u nano steps, %zu %s links: "
                "%ld of %ld cells differ from the whole-space run, %zu launches; %zu regions device to device (%zu across GPUs), "
                "%zu host pulls during the run\n", name, nodes, ghostZoneWidth, steps, links, LinkStats<Link>::name(), bad, cells, launches,
                direct, peer, pullsBefore);
}

template<typename CELL>
static void runBricks(const char *name, const Coord<3>& dim, int nodes, unsigned ghostZoneWidth, unsigned rounds)
{
    typedef typename B200Stepper<CELL>::GridType GridType;
    runBricksWith<CELL, PatchBuffer<GridType, GridType> >(name, dim, nodes, ghostZoneWidth, rounds);
    runBricksWith<CELL, B200PatchLink<CELL> >(name, dim, nodes, ghostZoneWidth, rounds);
}

/* the two ends of a B200PatchLink need not both be B200Steppers: a region put from a host grid is taken by a device
 * grid and the other way round (the bookkeeping of PatchBuffer either way) */
static void testMixedLinkEnds()
{
    typedef Jacobi7Cube CELL;
    typedef B200Stepper<CELL>::GridType HostGrid;
    const Coord<3> dim(12, 6, 5);
    CoordBox<3> box(Coord<3>(), dim);
    HostGrid host(box, CELL(0.5), CELL(0.25), dim);
    for (CoordBox<3>::Iterator i = box.begin(); i != box.end(); ++i) {
        host.set(*i, Seed<CELL>::make(i->toIndex(dim)));
    }
    B200Grid<CELL> dev(box, CELL(0.25));
    Region<3> region;
    region << CoordBox<3>(Coord<3>(2, 1, 1), Coord<3>(7, 3, 2)) << Streak<3>(Coord<3>(0, 5, 4), 12);
    B200PatchLink<CELL> link(region);
    link.pushRequest(4);
    link.pushRequest(8);
    link.put(host, region, dim, 3, 0);      /* not asked for: ignored */
    link.put(host, region, dim, 4, 0);
    CHECK(link.nextAvailableNanoStep() == 4 && link.nextRequiredNanoStep() == 8);
    link.getDevice(&dev, region, dim, 4, 0, true);
    long bad = 0;
    for (Region<3>::Iterator i = region.begin(); i != region.end(); ++i) {
        bad += !(dev.get(*i) == host.get(*i));
    }
    CHECK(bad == 0);
    CHECK(dev.get(Coord<3>(0, 0, 0)) == CELL());   /* nothing outside the region was touched */
    dev.set(Coord<3>(3, 2, 1), CELL(7.5));
    dev.set(Coord<3>(11, 5, 4), CELL(-2.25));
    link.putDevice(dev, region, dim, 8, 0);
    HostGrid back(box, CELL(0.5), CELL(0.25), dim);
    link.get(&back, region, dim, 8, 0, true);
    for (Region<3>::Iterator i = region.begin(); i != region.end(); ++i) {
        bad += !(back.get(*i) == dev.get(*i));
    }
    CHECK(bad == 0);
    CHECK(back.get(Coord<3>(3, 2, 1)) == CELL(7.5) && back.get(Coord<3>(0, 0, 0)) == CELL(0.5));
    CHECK(link.hostTransferCount() == 1 && link.deviceTransferCount() == 1);
    bool empty = false;
    try {
        link.getDevice(&dev, region, dim, 12, 0, true);
    } catch (const std::logic_error&) {
        empty = true;
    }
    CHECK(empty);
    std::printf("B200PatchLink with a host grid at one end: %ld cells differ\n", bad);
}

int main()
{
    try {
        for (unsigned width = 1; width <= 4; ++width) {
            runCase<ConwayCube>("ConwayCube", Coord<2>(17, 2 * width + 14), width, 9);   /* the reference test's space is 17 x 12 */
            runCase<Jacobi7Cube>("Jacobi7Cube", Coord<3>(11, 7, 2 * width + 13), width, 9);
        }
        runCase<Jacobi27Cube>("Jacobi27Cube", Coord<3>(9, 8, 19), 3, 8);
        runCase<ConwayCube>("ConwayCube", Coord<2>(40, 30), 2, 7);
        runBricks<Jacobi7Cube>("Jacobi7Cube", Coord<3>(20, 18, 16), 8, 2, 4);     /* 2 x 2 x 2 */
        runBricks<Jacobi27Cube>("Jacobi27Cube", Coord<3>(21, 17, 19), 8, 3, 3);   /* corner and edge neighbours matter */
        runBricks<Jacobi7Cube>("Jacobi7Cube", Coord<3>(24, 10, 9), 6, 1, 5);
        testMixedLinkEnds();
        /* a Torus cell is refused */
        bool refused = false;
        try {
            typedef Topologies::Torus<3>::Topology Topology;
            SharedPtr<PartitionManager<Topology> >::Type manager(new PartitionManager<Topology>());
            SharedPtr<SeededInitializer<Jacobi7Torus> >::Type init(new SeededInitializer<Jacobi7Torus>(Coord<3>(8, 8, 8), 2));
            SharedPtr<Partition<3> >::Type partition(new StripingPartition<3>(Coord<3>(), Coord<3>(8, 8, 8), 0, std::vector<std::size_t>(1, 512)));
            SharedPtr<AdjacencyManufacturer<3> >::Type adjacency(new DummyAdjacencyManufacturer<3>);
            manager->resetRegions(adjacency, CoordBox<3>(Coord<3>(), Coord<3>(8, 8, 8)), partition, 0, 1);
            manager->resetGhostZones(std::vector<CoordBox<3> >(1, CoordBox<3>(Coord<3>(), Coord<3>(8, 8, 8))),
                                     std::vector<CoordBox<3> >(1, CoordBox<3>(Coord<3>(), Coord<3>(8, 8, 8))));
            B200Stepper<Jacobi7Torus> stepper(manager, init);
        } catch (const std::logic_error&) {
            refused = true;
        }
        CHECK(refused);
    } catch (const std::exception& e) {
        std::printf("FAILED with exception: %s\n", e.what());
        return 2;
    }
    if (failures) {
        std::printf("%d check(s) FAILED\n", failures);
        return 1;
    }
    std::printf("stepper_test: all checks passed\n");
    return 0;
}
