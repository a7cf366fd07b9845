/* CPU test of the generic SoA path's HOST side and address arithmetic (include/libgeodecomp_b200/b200genericsoa.h):
 * unbound SoA cells (soa_cells.h) run through the reference's SerialSimulator and through B200Simulator /
 * B200StripingSimulator on the mock engine (mock_b200geo.cpp, uniform element layout on host memory). The device
 * kernel is replaced by HostSweep below — a plain loop over the box calling the SAME updateCell<CELL, STRIDE>() the
 * kernel calls per thread, through the SAME stride dispatch — so member-table probing, stride selection, the hood's
 * index arithmetic over the ghost ring (EDGE and WRAP layers) and the member-major cell streams are checked against
 * the reference bit for bit without a GPU. The launch itself is what generic_soa_test.cu covers on the device. */
#include <libgeodecomp/io/simpleinitializer.h>
#include <libgeodecomp/parallelization/serialsimulator.h>
#include <libgeodecomp/storage/soagrid.h>

#include <cstdio>
#include <cstring>
#include <vector>

#include "soa_cells.h"

#include <libgeodecomp_b200/b200stripingsimulator.h>

using namespace LibGeoDecomp;
using namespace soacells;

static int failures = 0;
#define CHECK(COND)                                                                     \
    do {                                                                                \
        if (!(COND)) {                                                                  \
            ++failures;                                                                 \
            std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #COND);              \
        }                                                                               \
    } while (0)

/* the mock engine's "device" memory is host memory: sweep the box with a loop */
struct HostSweep {
    template<typename CELL, long STRIDE>
    static void run(const B200Generic::SoA::BoxArgs& a)
    {
        ++launches;
        lastStride = STRIDE;
        for (int z = 0; z < a.dim[2]; ++z) {
            for (int y = 0; y < a.dim[1]; ++y) {
                for (int x = 0; x < a.dim[0]; ++x) {
                    B200Generic::SoA::updateCell<CELL, STRIDE>(
                        a.oldData, a.newData, a.first + x + y * a.pitch + z * a.plane, a.pitch, a.plane, a.nanoStep);
                }
            }
        }
    }

    static bool selectDevice(int)
    {
        return true;
    }

    static long launches;
    static long lastStride;
};
long HostSweep::launches = 0;
long HostSweep::lastStride = 0;

namespace LibGeoDecomp {
#define HOST_BINDING(CELL) \
    template<> struct B200KernelBinding<CELL> : public B200Generic::SoA::Binding<CELL, HostSweep> {};
HOST_BINDING(HeatSoACube)
HOST_BINDING(HeatSoATorus)
HOST_BINDING(MixSoACube)
HOST_BINDING(MixSoATorus)
}

static uint64_t splitmix(uint64_t x)
{
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

static double uniform(uint64_t i)
{
    return (double)(splitmix(i) >> 11) * (1.0 / 9007199254740992.0);
}

template<typename CELL> struct Seed;
template<typename T> struct Seed<HeatSoA<T> > {
    static HeatSoA<T> make(uint64_t i) { return HeatSoA<T>(uniform(i)); }
    static HeatSoA<T> edge() { return HeatSoA<T>(0.25); }
};
template<typename T> struct Seed<MixSoA<T> > {
    static MixSoA<T> make(uint64_t i)
    {
        MixSoA<T> c;
        c.density = uniform(5 * i);
        c.flux[0] = (float)uniform(5 * i + 1);
        c.flux[1] = (float)uniform(5 * i + 2);
        c.flux[2] = (float)uniform(5 * i + 3);
        c.count = (int)(splitmix(5 * i + 4) % 1000);
        c.tag = (short)(splitmix(7 * i) % 100);
        c.flag = (char)(splitmix(11 * i) % 2);
        return c;
    }
    static MixSoA<T> edge()
    {
        MixSoA<T> c;
        c.density = 0.5;
        c.flux[0] = 1.0f;
        c.flux[1] = 2.0f;
        c.flux[2] = 3.0f;
        c.count = 7;
        c.tag = 3;
        c.flag = 1;
        return c;
    }
};

template<typename CELL>
class SeededInitializer : public SimpleInitializer<CELL>
{
public:
    typedef typename SimpleInitializer<CELL>::Topology Topology;
    static const int DIM = Topology::DIM;
    using SimpleInitializer<CELL>::gridDimensions;

    SeededInitializer(const Coord<DIM>& dim, unsigned steps) : SimpleInitializer<CELL>(dim, steps) {}

    virtual void grid(GridBase<CELL, DIM> *ret)
    {
        CoordBox<DIM> box = ret->boundingBox();
        ret->setEdge(Seed<CELL>::edge());
        for (typename CoordBox<DIM>::Iterator i = box.begin(); i != box.end(); ++i) {
            ret->set(*i, Seed<CELL>::make(i->toIndex(gridDimensions())));
        }
    }
};

template<typename CELL, int DIM>
static long mismatches(const GridBase<CELL, DIM>& a, const GridBase<CELL, DIM>& b)
{
    long bad = 0;
    CoordBox<DIM> box = a.boundingBox();
    for (typename CoordBox<DIM>::Iterator i = box.begin(); i != box.end(); ++i) {
        bad += !(a.get(*i) == b.get(*i));
    }
    return bad;
}

/* SerialSimulator vs B200Simulator, and vs B200StripingSimulator on `slabs` slabs */
template<typename CELL, int DIM>
static void compare(const char *name, const Coord<DIM>& dim, unsigned steps, int slabs)
{
    SerialSimulator<CELL> ref(new SeededInitializer<CELL>(dim, steps));
    ref.run();
    long before = HostSweep::launches;
    {
        B200Simulator<CELL> sim(new SeededInitializer<CELL>(dim, steps));
        sim.run();
        long bad = mismatches<CELL, DIM>(*ref.getGrid(), *sim.getGrid());
        CHECK(bad == 0);
        CHECK(HostSweep::launches - before == (long)steps * APITraits::SelectNanoSteps<CELL>::VALUE);
        std::printf("%-14s %-16s %u steps, member stride %ld: %ld cells differ from SerialSimulator\n",
                    name, dim.toString().c_str(), steps, HostSweep::lastStride, bad);
    }
    if (slabs > 0) {
        /* slabs along z on a slab group: PEER ghost planes of the uniform-layout grids, the group calling the binding's
         * updateCallback for every slab (b200geo_group_step_with) */
        B200StripingSimulator<CELL> sim(new SeededInitializer<CELL>(dim, steps), std::vector<int>(slabs, 0), 1);
        sim.run();
        long bad = mismatches<CELL, DIM>(*ref.getGrid(), *sim.getGrid());
        CHECK(bad == 0);
        std::pair<unsigned long long, unsigned long long> st = sim.stripedGrid().exchangeStatistics();
        CHECK(slabs == 1 || st.first >= steps);
        std::printf("%-14s %-16s on %d slabs: %ld cells differ from SerialSimulator (%llu exchanges)\n",
                    name, dim.toString().c_str(), slabs, bad, st.first);
    }
}

template<typename CELL>
static void checkMembers(const char *name, std::size_t expectMembers, int expectBytes)
{
    std::vector<B200Member> m = B200KernelBinding<CELL>::members();
    int bytes = 0;
    for (std::size_t i = 0; i < m.size(); ++i) {
        bytes += m[i].bytes;
    }
    CHECK(m.size() == expectMembers);
    CHECK(bytes == expectBytes);
    CHECK((std::size_t)bytes == LibFlatArray::aggregated_member_size<CELL>::VALUE);
    std::printf("%-14s member table:", name);
    for (std::size_t i = 0; i < m.size(); ++i) {
        std::printf(" %d@%zu", m[i].bytes, m[i].offsetInCell);
    }
    std::printf("\n");
}

int main()
{
    /* the member table derived from the generated accessors: registration order, element widths, AoS offsets */
    checkMembers<HeatSoACube>("HeatSoACube", 1, 8);
    checkMembers<MixSoACube>("MixSoACube", 7, 8 + 12 + 4 + 2 + 1);
    {
        std::vector<B200Member> m = B200KernelBinding<MixSoACube>::members();
        CHECK(m[0].offsetInCell == offsetof(MixSoACube, density) && m[0].bytes == 8);
        CHECK(m[1].offsetInCell == offsetof(MixSoACube, flux) && m[1].bytes == 4);
        CHECK(m[2].offsetInCell == offsetof(MixSoACube, flux) + 4 && m[3].offsetInCell == offsetof(MixSoACube, flux) + 8);
        CHECK(m[4].offsetInCell == offsetof(MixSoACube, count) && m[4].bytes == 4);
        CHECK(m[5].offsetInCell == offsetof(MixSoACube, tag) && m[5].bytes == 2);
        CHECK(m[6].offsetInCell == offsetof(MixSoACube, flag) && m[6].bytes == 1);
    }

    /* stride list */
    CHECK(B200Generic::SoA::chooseStride(1) == (1L << B200GEO_SOA_STRIDE_MIN_LOG2));
    CHECK(B200Generic::SoA::chooseStride((1L << 20) + 1) == (3L << 19));
    CHECK(B200Generic::SoA::chooseStride((3L << 19) + 1) == (1L << 21));
    bool thrown = false;
    try {
        B200Generic::SoA::chooseStride((1L << B200GEO_SOA_STRIDE_MAX_LOG2) + 1);
    } catch (const std::out_of_range&) {
        thrown = true;
    }
    CHECK(thrown);

    compare<HeatSoACube, 3>("HeatSoACube", Coord<3>(20, 11, 7), 9, 1);
    compare<HeatSoATorus, 3>("HeatSoATorus", Coord<3>(20, 11, 7), 9, 2);
    compare<HeatSoACube, 3>("HeatSoACube", Coord<3>(70, 40, 33), 3, 0);      /* second stride of the list */
    compare<MixSoACube, 3>("MixSoACube", Coord<3>(13, 9, 6), 7, 3);
    compare<MixSoATorus, 3>("MixSoATorus", Coord<3>(13, 9, 6), 7, 2);

    std::printf(failures == 0 ? "generic SoA host test: all checks passed\n" : "generic SoA host test: %d FAILED\n", failures);
    return failures == 0 ? 0 : 1;
}
