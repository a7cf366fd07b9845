/* B200Stepper<CELL>: the Stepper of the reference's nesting layer on one B200.
 *
 * Plugs in where VanillaStepper / CUDAStepper do — as the STEPPER template argument of HiParSimulator
 * (parallelization/hiparsimulator.h:34-38) and UpdateGroup (parallelization/nesting/updategroup.h:28-29):
 *
 *     HiParSimulator<Cell, RecursiveBisectionPartition<3>, B200Stepper<Cell> > sim(...);
 *
 * Same constructor, same PatchAccepter / PatchProvider protocol (ghost zone links, writer and steerer adapters),
 * same schedule as VanillaStepper (parallelization/nesting/vanillastepper.h:93-225): ghost zones of width k are
 * computed k nano steps ahead in updateGhost() — so that the PatchLinks can ship them while the kernel is updated —
 * and the inner sets shrink step by step in update1() until the next synchronisation.
 *
 * What is different from the reference's CUDAStepper (parallelization/nesting/cudastepper.h:256-594), which keeps the
 * kernel on the device but updates the rims ON THE HOST and moves them streak by streak: here EVERYTHING lives on the
 * device. One device-resident double-buffered grid (B200Grid<CELL>) covers the bounding box of the node's expanded
 * region; inner sets AND rims are updated there, one launch per box of the region (B200Grid::updateRegion — regions
 * of a striping or bisection partition are a few boxes); the rim and volatile-kernel patch buffers
 * (commonstepper.h:236-283, storage/patchbufferfixed.h) are device memory, filled and drained by one region-copy
 * launch each. The host grid of CommonStepper only stages what PatchAccepters / PatchProviders see: a region is pulled
 * from the device right before an accepter that is due reads it, and pushed back right after a provider that is due
 * wrote it — nothing moves at the nano steps in between.
 *
 * Cube topologies (any partition); cells bound to a kernel family with B200GEO_BIND_CELL, or unbound cells in an nvcc
 * translation unit (generic device path). Needs no MPI: MPI enters only through the PatchLinks the caller adds.
 */
#ifndef LIBGEODECOMP_B200_B200STEPPER_H
#define LIBGEODECOMP_B200_B200STEPPER_H

#include <libgeodecomp/parallelization/nesting/commonstepper.h>
#include <libgeodecomp/storage/serializationbuffer.h>

#include "b200simulator.h"
#include "b200patchlink.h"

namespace LibGeoDecomp {

template<typename CELL_TYPE>
class B200Stepper : public CommonStepper<CELL_TYPE>
{
public:
    typedef typename Stepper<CELL_TYPE>::Topology Topology;
    const static int DIM = Topology::DIM;
    const static unsigned NANO_STEPS = APITraits::SelectNanoSteps<CELL_TYPE>::VALUE;

    typedef class CommonStepper<CELL_TYPE> ParentType;
    typedef typename ParentType::GridType GridType;
    typedef B200Grid<CELL_TYPE> DeviceGridType;
    typedef typename SerializationBuffer<CELL_TYPE>::BufferType BufferType;
    typedef typename ParentType::PatchAccepterVec PatchAccepterVec;
    typedef typename ParentType::PatchProviderVec PatchProviderVec;
    typedef typename ParentType::PatchAccepterList PatchAccepterList;
    typedef typename ParentType::PatchProviderList PatchProviderList;
    typedef typename ParentType::PatchType PatchType;
    typedef typename ParentType::InitPtr InitPtr;
    typedef typename ParentType::PartitionManagerPtr PartitionManagerPtr;

    using ParentType::initializer;
    using ParentType::patchAccepters;
    using ParentType::patchProviders;
    using ParentType::partitionManager;
    using ParentType::chronometer;
    using ParentType::innerSet;
    using ParentType::globalNanoStep;
    using ParentType::rim;
    using ParentType::resetValidGhostZoneWidth;
    using ParentType::initGridsCommon;
    using ParentType::getVolatileKernel;
    using ParentType::curStep;
    using ParentType::curNanoStep;
    using ParentType::validGhostZoneWidth;
    using ParentType::ghostZoneWidth;
    using ParentType::oldGrid;
    using ParentType::newGrid;

    /* the constructor of VanillaStepper / CUDAStepper, plus the device to run on */
    inline B200Stepper(
        PartitionManagerPtr partitionManager,
        InitPtr initializer,
        const PatchAccepterVec& ghostZonePatchAccepters = PatchAccepterVec(),
        const PatchAccepterVec& innerSetPatchAccepters = PatchAccepterVec(),
        const PatchProviderVec& ghostZonePatchProvidersPhase0 = PatchProviderVec(),
        const PatchProviderVec& ghostZonePatchProvidersPhase1 = PatchProviderVec(),
        const PatchProviderVec& innerSetPatchProviders = PatchProviderVec(),
        bool enableFineGrainedParallelism = false,
        int device = 0) :
        ParentType(
            partitionManager,
            initializer,
            ghostZonePatchAccepters,
            innerSetPatchAccepters,
            ghostZonePatchProvidersPhase0,
            ghostZonePatchProvidersPhase1,
            innerSetPatchProviders,
            enableFineGrainedParallelism),
        device(device),
        kernelPatch(0),
        hostIsCurrent(false),
        launches(0),
        pulls(0),
        pushes(0)
    {
        for (int i = 0; i < DIM; ++i) {
            if (Topology::wrapsAxis(i)) {
                throw std::logic_error("B200Stepper: Cube topologies only (a node's displaced grid may straddle a Torus seam)");
            }
        }
        rimPatch[0].data = rimPatch[1].data = 0;
        rimPatch[0].used = rimPatch[1].used = false;
        initGrids();
    }

    virtual ~B200Stepper()
    {
        b200geo_device_free(device, rimPatch[0].data);
        b200geo_device_free(device, rimPatch[1].data);
        b200geo_device_free(device, kernelPatch);
    }

    /* the whole grid as the host sees it: pulled from the device when it is asked for */
    virtual const GridType& grid() const
    {
        if (!hostIsCurrent) {
            pull(wholeBox);
            hostIsCurrent = true;
        }
        return *oldGrid;
    }

    const DeviceGridType& deviceGrid() const
    {
        return *deviceGridPtr;
    }

    /* kernel launches, device -> host and host -> device region transfers so far */
    std::size_t launchCount() const { return launches; }
    std::size_t pullCount() const { return pulls; }
    std::size_t pushCount() const { return pushes; }
    /* patches handed to / taken from accepters and providers that work on the device grid (B200PatchLink) */
    std::size_t devicePutCount() const { return devicePuts; }
    std::size_t deviceGetCount() const { return deviceGets; }

    /* Proceed the simulation exactly one nano step (vanillastepper.h:93-135) */
    virtual void update1()
    {
        TimeTotal t(&chronometer);
        unsigned index = ghostZoneWidth() - --validGhostZoneWidth;
        {
            TimeComputeInner timer(&chronometer);
            launches += deviceGridPtr->updateRegion(innerSet(index), curNanoStep);
            deviceGridPtr->swapBuffers();
            hostIsCurrent = false;
            lastPulledValid = false;

            ++curNanoStep;
            if (curNanoStep == NANO_STEPS) {
                curNanoStep = 0;
                ++curStep;
            }
        }

        notifyAccepters(innerSet(ghostZoneWidth()), ParentType::INNER_SET, globalNanoStep());

        if (validGhostZoneWidth == 0) {
            updateGhost();
            resetValidGhostZoneWidth();
        }

        index = ghostZoneWidth() - validGhostZoneWidth;
        notifyProviders(innerSet(index), ParentType::INNER_SET, globalNanoStep());
    }

private:
    struct DevicePatch {
        void *data;
        std::size_t nanoStep;
        bool used;
    };

    int device;
    typename SharedPtr<DeviceGridType>::Type deviceGridPtr;
    Region<DIM> wholeBox;
    DevicePatch rimPatch[2];      /* PatchBufferFixed<GridType, GridType, 2> rimBuffer, on the device */
    void *kernelPatch;            /* PatchBufferFixed<GridType, GridType, 1> kernelBuffer, on the device */
    mutable bool hostIsCurrent;
    std::size_t launches;
    mutable std::size_t pulls;
    std::size_t pushes;
    std::size_t devicePuts = 0, deviceGets = 0;
    Region<DIM> lastPulled;
    std::size_t lastPulledNanoStep = 0;
    bool lastPulledValid = false;

    /* device (current buffer) -> host grid */
    void pull(const Region<DIM>& region) const
    {
        if (region.empty()) {
            return;
        }
        BufferType buffer;
        deviceGridPtr->saveRegion(&buffer, region);
        oldGrid->loadRegion(buffer, region);
        ++pulls;
    }

    /* host grid -> device (current buffer only: the two buffers differ on purpose, as oldGrid and newGrid do) */
    void push(const GridType& from, const Region<DIM>& region)
    {
        if (region.empty()) {
            return;
        }
        BufferType buffer = SerializationBuffer<CELL_TYPE>::create(region);
        from.saveRegion(&buffer, region);
        pushBuffer(buffer, region);
        ++pushes;
    }

    void pushBuffer(const std::vector<char>& buffer, const Region<DIM>& region)
    {
        deviceGridPtr->loadRegionBytes(buffer, region, Coord<DIM>(), 0);
    }

    void pushBuffer(const std::vector<CELL_TYPE>& buffer, const Region<DIM>& region)
    {
        deviceGridPtr->loadRegionCells(buffer, region, Coord<DIM>(), 0);
    }

    /* PatchAccepters read the HOST grid: bring the region over only if one of them is due at this nano step. Accepters
     * that take the device grid (B200DevicePatchAccepter) get it as it is. Same order, same conditions as
     * CommonStepper::notifyPatchAccepters (commonstepper.h:112-131). */
    void notifyAccepters(const Region<DIM>& region, const PatchType& patchType, std::size_t nanoStep)
    {
        TimePatchAccepters t(&chronometer);
        for (typename PatchAccepterList::iterator i = patchAccepters[patchType].begin(); i != patchAccepters[patchType].end(); ++i) {
            if (nanoStep != (*i)->nextRequiredNanoStep()) {
                continue;
            }
            B200DevicePatchAccepter<CELL_TYPE> *direct = dynamic_cast<B200DevicePatchAccepter<CELL_TYPE>*>(&**i);
            if (direct) {
                direct->putDevice(*deviceGridPtr, region, partitionManager->getSimulationArea(), nanoStep, partitionManager->rank());
                ++devicePuts;
                continue;
            }
            if (!hostIsCurrent && !pulledFor(region, nanoStep)) {
                pull(region);
                lastPulled = region;
                lastPulledNanoStep = nanoStep;
                lastPulledValid = true;
            }
            (*i)->put(*oldGrid, region, partitionManager->getSimulationArea(), nanoStep, partitionManager->rank());
        }
    }

    /* PatchProviders write the HOST grid: hand them the current cells, take the region back afterwards; providers that
     * take the device grid (B200DevicePatchProvider) write it directly (commonstepper.h:133-153) */
    void notifyProviders(const Region<DIM>& region, const PatchType& patchType, std::size_t nanoStep)
    {
        TimePatchProviders t(&chronometer);
        for (typename PatchProviderList::iterator i = patchProviders[patchType].begin(); i != patchProviders[patchType].end(); ++i) {
            if (nanoStep != (*i)->nextAvailableNanoStep()) {
                continue;
            }
            B200DevicePatchProvider<CELL_TYPE> *direct = dynamic_cast<B200DevicePatchProvider<CELL_TYPE>*>(&**i);
            if (direct) {
                direct->getDevice(&*deviceGridPtr, region, partitionManager->getSimulationArea(), nanoStep, partitionManager->rank(), true);
                hostIsCurrent = false;
                lastPulledValid = false;
                ++deviceGets;
                continue;
            }
            if (!hostIsCurrent) {
                pull(region);
            }
            (*i)->get(&*oldGrid, region, partitionManager->getSimulationArea(), nanoStep, partitionManager->rank(), true);
            push(*oldGrid, region);
            lastPulledValid = false;
        }
    }

    /* several host-side accepters due at the same nano step for the same region share one pull */
    bool pulledFor(const Region<DIM>& region, std::size_t nanoStep) const
    {
        return lastPulledValid && lastPulledNanoStep == nanoStep && lastPulled == region;
    }

    void saveRim(std::size_t nanoStep)
    {
        for (int i = 0; i < 2; ++i) {
            if (!rimPatch[i].used) {
                deviceGridPtr->saveRegionToDevice(rimPatch[i].data, rim());
                rimPatch[i].nanoStep = nanoStep;
                rimPatch[i].used = true;
                return;
            }
        }
        throw std::logic_error("B200Stepper: rim buffer full");
    }

    void restoreRim(bool remove)
    {
        for (int i = 0; i < 2; ++i) {
            if (rimPatch[i].used && rimPatch[i].nanoStep == globalNanoStep()) {
                deviceGridPtr->loadRegionFromDevice(rimPatch[i].data, rim());
                hostIsCurrent = false;
                lastPulledValid = false;
                if (remove) {
                    rimPatch[i].used = false;
                }
                return;
            }
        }
        throw std::logic_error("B200Stepper: no rim stored for this nano step");
    }

    void saveKernel()
    {
        deviceGridPtr->saveRegionToDevice(kernelPatch, getVolatileKernel());
    }

    void restoreKernel()
    {
        deviceGridPtr->loadRegionFromDevice(kernelPatch, getVolatileKernel());
        hostIsCurrent = false;
        lastPulledValid = false;
    }

    inline void initGrids()
    {
        CoordBox<DIM> gridBox = initGridsCommon();
        wholeBox.clear();
        wholeBox << gridBox;

        deviceGridPtr.reset(new DeviceGridType(gridBox, oldGrid->getEdge(), device));
        /* scratch buffer = newGrid, current buffer = oldGrid (commonstepper.h:158-169: the providers of step 0
         * have written oldGrid only) */
        push(*newGrid, wholeBox);
        deviceGridPtr->swapBuffers();
        push(*oldGrid, wholeBox);
        hostIsCurrent = true;

        const std::size_t cellBytes = (std::size_t)deviceGridPtr->bytesPerCell();
        for (int i = 0; i < 2; ++i) {
            B200Helpers::check(b200geo_device_alloc(device, rim().size() * cellBytes, &rimPatch[i].data));
        }
        B200Helpers::check(b200geo_device_alloc(device, getVolatileKernel().size() * cellBytes, &kernelPatch));

        notifyAccepters(rim(), ParentType::GHOST_PHASE_0, globalNanoStep());
        notifyAccepters(innerSet(ghostZoneWidth()), ParentType::INNER_SET, globalNanoStep());

        saveRim(globalNanoStep());
        updateGhost();
    }

    /* computes the next ghost zone, ghostZoneWidth() nano steps ahead (vanillastepper.h:157-225), on the device */
    inline void updateGhost()
    {
        {
            TimeComputeGhost t(&chronometer);
            /* the ghost zone update destroys parts of the kernel: save them; the rim was destroyed while the
             * kernel was updated: restore it */
            saveKernel();
            restoreRim(false);
        }

        std::size_t oldNanoStep = curNanoStep;
        std::size_t oldStep = curStep;
        std::size_t curGlobalNanoStep = globalNanoStep();

        for (std::size_t t = 0; t < ghostZoneWidth(); ++t) {
            notifyProviders(rim(t), ParentType::GHOST_PHASE_0, globalNanoStep());
            notifyProviders(rim(t), ParentType::GHOST_PHASE_1, globalNanoStep());

            {
                TimeComputeGhost timer(&chronometer);
                launches += deviceGridPtr->updateRegion(rim(t + 1), curNanoStep);

                ++curNanoStep;
                if (curNanoStep == NANO_STEPS) {
                    curNanoStep = 0;
                    curStep++;
                }

                deviceGridPtr->swapBuffers();
                hostIsCurrent = false;
                lastPulledValid = false;
                ++curGlobalNanoStep;
            }

            notifyAccepters(rim(ghostZoneWidth()), ParentType::GHOST_PHASE_0, curGlobalNanoStep);
        }

        {
            TimeComputeGhost t(&chronometer);
            saveRim(curGlobalNanoStep);
            if (ghostZoneWidth() % 2) {
                deviceGridPtr->swapBuffers();
            }

            /* back to the kernel's time */
            curNanoStep = oldNanoStep;
            curStep = oldStep;
            restoreRim(true);
            restoreKernel();
        }
    }
};

}

#endif
