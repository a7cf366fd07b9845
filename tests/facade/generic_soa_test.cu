/* Drop-in test of the GENERIC SoA device path (include/libgeodecomp_b200/b200genericsoa.h): Struct-of-Arrays user
 * cells WITHOUT a B200GEO_BIND_CELL line (soa_cells.h), their own SoA-signature updateLineX() compiled by nvcc into
 * sm_100a kernels that use the accessors LIBFLATARRAY_REGISTER_SOA generated. Every case runs the reference's
 * SerialSimulator (SoAGrid + FixedNeighborhoodUpdateFunctor) beside B200Simulator / B200StripingSimulator in this
 * process and requires bit-identical grids. The host side of the same path (member table, stride dispatch, hood
 * arithmetic) is covered on the CPU by generic_soa_host_test.cpp. Compiled HERE (tests/facade/Makefile), run on the
 * GPU box by tests/test_facade_gpu.py. `--bench` prints the throughput of the heat cell. */
#include <cuda.h>

#include <libgeodecomp/io/simpleinitializer.h>
#include <libgeodecomp/parallelization/serialsimulator.h>
#include <libgeodecomp/storage/soagrid.h>

#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "soa_cells.h"

#include <libgeodecomp_b200/b200stripingsimulator.h>

using namespace LibGeoDecomp;
using namespace soacells;

/* which generic path a cell takes is decided at compile time */
#include <libgeodecomp/misc/testcell.h>
#include <type_traits>
static_assert(std::is_base_of<B200Generic::SoA::Binding<HeatSoACube, B200Generic::SoA::DeviceSweep>, B200KernelBinding<HeatSoACube> >::value,
              "a SoA cell whose only update is a SoA-signature updateLineX() takes the SoA path");
static_assert(std::is_base_of<B200Generic::SoA::Binding<MixSoATorus, B200Generic::SoA::DeviceSweep>, B200KernelBinding<MixSoATorus> >::value,
              "a SoA cell whose only update is a SoA-signature updateLineX() takes the SoA path");
static_assert(std::is_base_of<B200Generic::WordSlicedBinding<TestCellSoA>, B200KernelBinding<TestCellSoA> >::value,
              "a SoA cell with a per-cell update() stays word-sliced: the reference's TestCellSoA");
static_assert(std::is_base_of<B200Generic::WordSlicedBinding<TestCell<3> >, B200KernelBinding<TestCell<3> > >::value,
              "AoS cells are word-sliced");

static int failures = 0;
#define CHECK(COND)                                                                     \
    do {                                                                                \
        if (!(COND)) {                                                                  \
            ++failures;                                                                 \
            std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #COND);              \
        }                                                                               \
    } while (0)

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

    /* rows through set(Streak, cells): one transfer per row */
    virtual void grid(GridBase<CELL, DIM> *ret)
    {
        CoordBox<DIM> box = ret->boundingBox();
        ret->setEdge(Seed<CELL>::edge());
        std::vector<CELL> row(box.dimensions.x());
        for (int z = box.origin.z(); z < box.origin.z() + box.dimensions.z(); ++z) {
            for (int y = box.origin.y(); y < box.origin.y() + box.dimensions.y(); ++y) {
                for (int x = 0; x < box.dimensions.x(); ++x) {
                    row[x] = Seed<CELL>::make(Coord<3>(box.origin.x() + x, y, z).toIndex(gridDimensions()));
                }
                ret->set(Streak<3>(Coord<3>(box.origin.x(), y, z), box.origin.x() + box.dimensions.x()), row.data());
            }
        }
    }
};

template<typename CELL>
static long differingCells(const GridBase<CELL, 3> *a, const GridBase<CELL, 3> *b)
{
    long bad = 0;
    CoordBox<3> box = a->boundingBox();
    std::vector<CELL> rowA(box.dimensions.x()), rowB(box.dimensions.x());
    for (int z = 0; z < box.dimensions.z(); ++z) {
        for (int y = 0; y < box.dimensions.y(); ++y) {
            Streak<3> s(Coord<3>(box.origin.x(), box.origin.y() + y, box.origin.z() + z), box.origin.x() + box.dimensions.x());
            a->get(s, rowA.data());
            b->get(s, rowB.data());
            for (int x = 0; x < box.dimensions.x(); ++x) {
                bad += !(rowA[x] == rowB[x]);
            }
        }
    }
    return bad;
}

static std::vector<int> devicesFor(int slabs)
{
    int n = b200geo_device_count();
    std::vector<int> ret;
    for (int i = 0; i < slabs; ++i) {
        ret.push_back(n > 0 ? i % n : 0);
    }
    return ret;
}

template<typename CELL>
static void compareWithSerialSimulator(const char *name, const Coord<3>& dim, unsigned steps)
{
    SerialSimulator<CELL> ref(new SeededInitializer<CELL>(dim, steps));
    B200Simulator<CELL> sim(new SeededInitializer<CELL>(dim, steps));
    ref.run();
    sim.run();
    CHECK(sim.getStep() == steps);
    long bad = differingCells<CELL>(ref.getGrid(), sim.getGrid());
    CHECK(bad == 0);
    std::printf("%-14s %-16s: %s (%ld differing cells after %u steps)\n", name, dim.toString().c_str(),
                bad ? "MISMATCH" : "bit-exact", bad, steps);
}

template<typename CELL>
static void compareStriped(const char *name, const Coord<3>& dim, unsigned steps, int slabs)
{
    SerialSimulator<CELL> ref(new SeededInitializer<CELL>(dim, steps));
    B200StripingSimulator<CELL> sim(new SeededInitializer<CELL>(dim, steps), devicesFor(slabs));
    ref.run();
    sim.run();
    CHECK(sim.getStep() == steps);
    long bad = differingCells<CELL>(ref.getGrid(), sim.getGrid());
    CHECK(bad == 0);
    std::pair<unsigned long long, unsigned long long> st = sim.stripedGrid().exchangeStatistics();
    CHECK(slabs == 1 || st.first >= steps);
    std::printf("%-14s on %d slabs: %s (%ld differing cells after %u steps, %llu exchanges)\n", name, slabs,
                bad ? "MISMATCH" : "bit-exact", bad, steps, st.first);
}

template<typename CELL>
static void bench(const char *what, const Coord<3>& dim, unsigned steps, double bytesPerUpdate)
{
    B200Simulator<CELL> sim(new SeededInitializer<CELL>(dim, steps));
    sim.run();   // warm-up: first launch, module load
    auto t0 = std::chrono::steady_clock::now();
    sim.run();
    sim.getGrid();
    double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    double updates = (double)dim.prod() * steps * APITraits::SelectNanoSteps<CELL>::VALUE;
    std::printf("{\"workload\": \"%s\", \"glups\": %.2f, \"algorithmic_gbs\": %.1f, \"seconds\": %.4f, \"note\": \"run() incl. Initializer, wall clock\"}\n",
                what, updates / s * 1e-9, updates * bytesPerUpdate / s * 1e-9, s);
}

int main(int argc, char **argv)
{
    if (argc > 1 && std::string(argv[1]) == "--bench") {
        try {
#ifndef B200GEO_SOA_BENCH_ONLY
            std::printf("{\"note\": \"this binary lists member strides up to 2^22 elements only; the 384^3 bench is _bin/generic_soa_bench\"}\n");
            return 0;
#endif
            bench<HeatSoACube>("HeatSoACube 7-point f64 384^3 (user SoA updateLineX, generated accessors)", Coord<3>(384, 384, 384), 200, 16);
        } catch (const std::exception& e) {
            std::printf("FAILED with exception: %s\n", e.what());
            return 2;
        }
        return 0;
    }
#ifdef B200GEO_SOA_BENCH_ONLY
    std::printf("built as generic_soa_bench: run with --bench\n");
    return 1;
#else
    try {
        CHECK(B200KernelBinding<HeatSoACube>::kernel() == B200GEO_KERNEL_GENERIC);
        CHECK(B200KernelBinding<HeatSoACube>::members().size() == 1);
        CHECK(B200KernelBinding<MixSoATorus>::members().size() == 7);

        compareWithSerialSimulator<HeatSoACube>("HeatSoACube", Coord<3>(20, 11, 7), 9);
        compareWithSerialSimulator<HeatSoATorus>("HeatSoATorus", Coord<3>(20, 11, 7), 9);
        compareWithSerialSimulator<HeatSoACube>("HeatSoACube", Coord<3>(130, 67, 33), 5);
        compareWithSerialSimulator<HeatSoATorus>("HeatSoATorus", Coord<3>(200, 40, 33), 4);
        compareWithSerialSimulator<MixSoACube>("MixSoACube", Coord<3>(13, 9, 6), 7);
        compareWithSerialSimulator<MixSoATorus>("MixSoATorus", Coord<3>(13, 9, 6), 7);
        compareWithSerialSimulator<MixSoACube>("MixSoACube", Coord<3>(150, 31, 12), 4);

        compareStriped<HeatSoACube>("HeatSoACube", Coord<3>(33, 10, 9), 9, 1);
        compareStriped<HeatSoACube>("HeatSoACube", Coord<3>(33, 10, 12), 9, 3);
        compareStriped<HeatSoATorus>("HeatSoATorus", Coord<3>(20, 11, 12), 6, 2);
        compareStriped<MixSoACube>("MixSoACube", Coord<3>(13, 9, 12), 7, 4);
        compareStriped<MixSoATorus>("MixSoATorus", Coord<3>(13, 9, 9), 5, 3);
    } catch (const std::exception& e) {
        std::printf("FAILED with exception: %s\n", e.what());
        return 2;
    }
    if (failures) {
        std::printf("%d check(s) FAILED\n", failures);
        return 1;
    }
    std::printf("generic_soa_test: all checks passed\n");
    return 0;
#endif
}
