/* End to end through the C++ façade — the reference's own host language: a LibGeoDecomp program (unchanged user cell,
 * an Initializer, a Writer, Simulator::run()) on B200Simulator / B200StripingSimulator, timed around run(), host
 * memory -> device -> host memory inside the timed region. What bench.py's e2e number is for the Python mirror, this is
 * for the C++ side (bench.py runs it and puts the line into "e2e_cpp").
 *
 *   e2e_bench [n = 1024] [steps = 20] [slabs = 1] [mode = box | rows | stream] [--bov prefix]
 *
 * mode stream: as box, but the result goes to a ParallelWriter (io/parallelwriter.h) instead of a serial Writer: with
 *            nothing but ParallelWriters registered B200Simulator::run() takes the streamed schedule
 *            (b200streamedrun.h): Initializer, sweeps and writer work on different chunks of the grid at the same time;
 * mode box : the Initializer hands the engine its whole box with GridBase::loadMember, the Writer pulls the final grid
 *            with GridBase::saveMember (storage/gridbase.h:217-261) — one strided copy per box, page-locked host memory
 *            (b200geo_host_alloc);
 * mode rows: the Initializer writes the grid row by row with GridBase::set(Streak, cells) (what SimpleInitializer-style
 *            user code does), the Writer pulls it row by row with saveMemberUnchecked — the access pattern of the
 *            reference's BOVOutput::writeGrid (io/bovoutput.h:83-95); the façade combines the writes and reads ahead.
 * --bov    : additionally a (smaller) run with the reference's unmodified SerialBOVWriter writing <prefix>.*.data.
 *
 * The grid is checked after the run: a position-weighted checksum of the pulled result against the same run through
 * bench's other modes is printed; parity itself is the business of the test binaries beside this one. */
#include <libgeodecomp/io/parallelwriter.h>
#include <libgeodecomp/io/serialbovwriter.h>
#include <libgeodecomp/io/simpleinitializer.h>
#include <libgeodecomp/misc/clonable.h>

#include <libgeodecomp_b200/b200stripingsimulator.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "bindings.h"

using namespace LibGeoDecomp;
using namespace b200models;

typedef Jacobi27Cube Cell;

static uint64_t splitmix(uint64_t x)
{
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

struct HostField {
    double *data;
    Coord<3> dim;
};

class BoxInitializer : public SimpleInitializer<Cell>
{
public:
    BoxInitializer(const HostField& field, unsigned steps, bool rows) :
        SimpleInitializer<Cell>(field.dim, steps), field(field), rows(rows) {}

    virtual void grid(GridBase<Cell, 3> *target)
    {
        CoordBox<3> box = target->boundingBox();
        if (rows) {
            for (CoordBox<3>::StreakIterator i = box.beginStreak(); i != box.endStreak(); ++i) {
                const double *row = field.data + ((std::size_t)i->origin.z() * field.dim.y() + i->origin.y()) * field.dim.x() + i->origin.x();
                target->set(*i, reinterpret_cast<const Cell*>(row));      /* a Jacobi cell is its one double */
            }
            return;
        }
        if (!(box == cachedBox)) {
            /* grid() is called by the simulator's constructor and again by run(): the Region (a run-length list of
             * 2^20 streaks for 1024^3) is built once */
            cachedRegion.clear();
            cachedRegion << box;
            cachedBox = box;
        }
        /* the box is contiguous in the host array when it spans whole planes (slabs along z do) */
        const double *first = field.data + (std::size_t)box.origin.z() * field.dim.y() * field.dim.x();
        target->loadMember(first, MemoryLocation::HOST, Selector<Cell>(&Cell::temp, "temp"), cachedRegion);
    }

private:
    HostField field;
    bool rows;
    CoordBox<3> cachedBox;
    Region<3> cachedRegion;
};

class PullWriter : public Clonable<Writer<Cell>, PullWriter>
{
public:
    PullWriter(const HostField& field, bool rows) : Clonable<Writer<Cell>, PullWriter>("", 1u << 30), field(field), rows(rows) {}

    virtual void stepFinished(const GridType& grid, unsigned, WriterEvent event)
    {
        if (event != WRITER_ALL_DONE) {
            return;
        }
        Selector<Cell> selector(&Cell::temp, "temp");
        CoordBox<3> box = grid.boundingBox();
        if (rows) {
            for (CoordBox<3>::StreakIterator i = box.beginStreak(); i != box.endStreak(); ++i) {
                Region<3> one;
                one << *i;
                double *row = field.data + ((std::size_t)i->origin.z() * field.dim.y() + i->origin.y()) * field.dim.x() + i->origin.x();
                grid.saveMemberUnchecked(reinterpret_cast<char*>(row), MemoryLocation::HOST, selector, one);
            }
            return;
        }
        if (!(box == cachedBox)) {
            cachedRegion.clear();
            cachedRegion << box;
            cachedBox = box;
        }
        grid.saveMember(field.data, MemoryLocation::HOST, selector, cachedRegion);
    }

    /* the Region of the whole grid is built ahead of the run (2^20 streaks for 1024^3) */
    void prepare(const CoordBox<3>& box)
    {
        cachedRegion.clear();
        cachedRegion << box;
        cachedBox = box;
    }

private:
    HostField field;
    bool rows;
    CoordBox<3> cachedBox;
    Region<3> cachedRegion;
};

/* the same writer under the reference's ParallelWriter interface: pulls whatever validRegion it is handed */
class ParallelPullWriter : public ParallelWriter<Cell>
{
public:
    explicit ParallelPullWriter(const HostField& field) : ParallelWriter<Cell>("", 1u << 30), field(field) {}

    virtual ParallelWriter<Cell> *clone() const
    {
        return new ParallelPullWriter(*this);
    }

    virtual void stepFinished(const GridType& grid, const RegionType& validRegion, const CoordType&, unsigned, WriterEvent event,
                              std::size_t, bool)
    {
        if (event != WRITER_ALL_DONE) {
            return;
        }
        /* whole planes (the simulator cuts along z): contiguous in the host array */
        CoordBox<3> box = validRegion.boundingBox();
        double *first = field.data + (std::size_t)box.origin.z() * field.dim.y() * field.dim.x();
        grid.saveMember(first, MemoryLocation::HOST, Selector<Cell>(&Cell::temp, "temp"), validRegion);
    }

private:
    HostField field;
};

static void fill(const HostField& f)
{
    /* a 32-plane block of noise repeated along z (generation speed), as bench.py does */
    const std::size_t plane = (std::size_t)f.dim.x() * f.dim.y();
    const int tile = f.dim.z() < 32 ? f.dim.z() : 32;
#pragma omp parallel for
    for (long long i = 0; i < (long long)(plane * tile); ++i) {
        f.data[i] = (double)(splitmix((uint64_t)i ^ 42) >> 11) * (1.0 / 9007199254740992.0);
    }
    for (int z = tile; z < f.dim.z(); z += tile) {
        int n = f.dim.z() - z < tile ? f.dim.z() - z : tile;
        std::memcpy(f.data + (std::size_t)z * plane, f.data, (std::size_t)n * plane * sizeof(double));
    }
}

static unsigned long long checksum(const HostField& f)
{
    /* over the bit patterns, position-weighted (swapped planes or rows do not cancel), in integer arithmetic
     * (the same value whatever the thread count) */
    const std::size_t n = (std::size_t)f.dim.prod();
    unsigned long long s = 0;
#pragma omp parallel for reduction(+ : s)
    for (long long i = 0; i < (long long)n; ++i) {
        unsigned long long bits;
        std::memcpy(&bits, &f.data[i], sizeof(bits));
        s += bits * (unsigned long long)(1 + (i % 1021));
    }
    return s;
}

template<typename SIM>
static void timeRun(SIM& sim, const HostField& field, const char *what, int n, unsigned steps, int slabs, const char *mode)
{
    if (std::string(mode) == "stream") {
        sim.addWriter(new ParallelPullWriter(field));
    } else {
        PullWriter *writer = new PullWriter(field, std::string(mode) == "rows");
        writer->prepare(CoordBox<3>(Coord<3>(), field.dim));
        sim.addWriter(writer);
    }
    auto t0 = std::chrono::steady_clock::now();
    sim.run();
    double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    double updates = (double)n * n * n * steps;
    std::printf("{\"impl\": \"%s\", \"cell\": \"Jacobi27Cube\", \"dims\": [%d, %d, %d], \"steps\": %u, \"slabs\": %d, \"mode\": \"%s\", "
                "\"seconds\": %.4f, \"value\": %.6f, \"unit\": \"GLUPS\", \"h2d_bytes_per_step\": %.0f, \"d2h_bytes_per_step\": %.0f, "
                "\"checksum\": \"%llu\", \"what\": \"C++ program: run() of the facade simulator, Initializer from / Writer into host memory, wall clock\"}\n",
                what, n, n, n, steps, slabs, mode, s, updates / s * 1e-9, (double)n * n * n * 8 / steps, (double)n * n * n * 8 / steps,
                checksum(field));
    std::fflush(stdout);
}

int main(int argc, char **argv)
{
    int n = argc > 1 ? std::atoi(argv[1]) : 1024;
    unsigned steps = argc > 2 ? (unsigned)std::atoi(argv[2]) : 20;
    int slabs = argc > 3 ? std::atoi(argv[3]) : 1;
    std::string mode = argc > 4 ? argv[4] : "box";
    std::string bov;
    for (int i = 5; i + 1 < argc; ++i) {
        if (std::string(argv[i]) == "--bov") {
            bov = argv[i + 1];
        }
    }
    try {
        HostField field;
        field.dim = Coord<3>(n, n, n);
        void *p = 0;
        B200Helpers::check(b200geo_host_alloc((uint64_t)n * n * n * sizeof(double), &p));
        field.data = static_cast<double*>(p);
        fill(field);
        const bool rows = mode == "rows";
        if (slabs <= 1) {
            B200Simulator<Cell> sim(new BoxInitializer(field, steps, rows));
            timeRun(sim, field, "B200Simulator", n, steps, 1, mode.c_str());
        } else {
            std::vector<int> devices;
            int count = b200geo_device_count();
            for (int i = 0; i < slabs; ++i) {
                devices.push_back(count > 0 ? i % count : 0);
            }
            B200StripingSimulator<Cell> sim(new BoxInitializer(field, steps, rows), devices);
            timeRun(sim, field, "B200StripingSimulator", n, steps, slabs, mode.c_str());
        }
        if (!bov.empty()) {
            /* the reference's own writer, unmodified: one brick per call, pulled row by row (io/bovoutput.h:83-95) */
            const int m = n < 256 ? n : 256;
            HostField small = field;
            small.dim = Coord<3>(m, m, m);
            B200Simulator<Cell> sim(new BoxInitializer(small, 4, false));
            sim.addWriter(new SerialBOVWriter<Cell>(Selector<Cell>(&Cell::temp, "temp"), bov, 4));
            auto t0 = std::chrono::steady_clock::now();
            sim.run();
            double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            std::printf("{\"impl\": \"B200Simulator + the reference's SerialBOVWriter\", \"dims\": [%d, %d, %d], \"steps\": 4, \"bricks\": 2, "
                        "\"seconds\": %.4f, \"brick_bytes\": %.0f}\n", m, m, m, s, (double)m * m * m * 8);
        }
        b200geo_host_free(field.data);
    } catch (const std::exception& e) {
        std::printf("{\"error\": \"%s\"}\n", e.what());
        return 2;
    }
    return 0;
}
