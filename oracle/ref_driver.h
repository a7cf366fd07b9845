/* TEST INFRASTRUCTURE — generic driver that runs a user model through the REFERENCE's own
 * SerialSimulator / OpenMPSimulator (parallelization/serialsimulator.h:23,
 * parallelization/openmpsimulator.h:20) on a grid read from a raw member-major file and
 * dumps the final grid in the same format. Compiled against the headers where they lie
 * under /root/reference by oracle/Makefile; binaries land in oracle/_ref/.
 *
 * Raw format ("member-major", identical to SoAGrid::saveRegion over the whole box,
 * storage/soagrid.h:523-576): for every member in registration order one dense array
 * [nz][ny][nx] (x fastest) of that member's type, little endian, no padding.
 *
 * usage: lgd_ref_<family> <model> <nx> <ny> <nz> <steps> <in.raw> <out.raw> [--omp] [--edge <v>] [--tile <t>]
 * --tile t: in.raw holds only t planes (nx x ny x t, member-major) and is repeated along z — how bench.py's synthetic
 * Jacobi input is built — so that a full-size run (1024^3) needs no full-size file; out.raw "-" = no output pass.
 * prints one JSON line with the reference's own TimeCompute interval (misc/chronometer.h).
 */
#ifndef B200GEO_ORACLE_REF_DRIVER_H
#define B200GEO_ORACLE_REF_DRIVER_H

#include <libgeodecomp/io/simpleinitializer.h>
#include <libgeodecomp/parallelization/serialsimulator.h>
#ifdef _OPENMP
#include <libgeodecomp/parallelization/openmpsimulator.h>
#include <omp.h>
#endif

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace refdriver {

using namespace LibGeoDecomp;

/* Per-model (de)serialisation: specialised next to each model. */
template<typename CELL> struct Codec;

template<int DIM> struct Dims;
template<> struct Dims<2> {
    static Coord<2> make(int nx, int ny, int) { return Coord<2>(nx, ny); }
    static Coord<2> rowOrigin(int y, int) { return Coord<2>(0, y); }
    static int ny(const Coord<2>& c) { return c.y(); }
    static int nz(const Coord<2>&) { return 1; }
};
template<> struct Dims<3> {
    static Coord<3> make(int nx, int ny, int nz) { return Coord<3>(nx, ny, nz); }
    static Coord<3> rowOrigin(int y, int z) { return Coord<3>(0, y, z); }
    static int ny(const Coord<3>& c) { return c.y(); }
    static int nz(const Coord<3>& c) { return c.z(); }
};

template<typename CELL>
class RawInitializer : public SimpleInitializer<CELL>
{
public:
    typedef typename SimpleInitializer<CELL>::Topology Topology;
    static const int DIM = Topology::DIM;

    RawInitializer(const Coord<DIM>& dim, unsigned steps, const std::vector<char> *raw, bool haveEdge, double edge, int tile = 0) :
        SimpleInitializer<CELL>(dim, steps), raw(raw), haveEdge(haveEdge), edge(edge), tile(tile)
    {}

    virtual void grid(GridBase<CELL, DIM> *ret)
    {
        Coord<DIM> dim = this->gridDimensions();
        CoordBox<DIM> box = ret->boundingBox();
        if (haveEdge) {
            ret->setEdge(Codec<CELL>::edge(edge));
        }
        std::size_t cells = (std::size_t)dim.prod();
        int nx = dim.x();
        if (tile > 0) {
            // the file holds `tile` planes, repeated along z
            cells = (std::size_t)nx * Dims<DIM>::ny(dim) * tile;
        }
        std::vector<CELL> row(nx);
        for (int z = 0; z < Dims<DIM>::nz(dim); ++z) {
            for (int y = 0; y < Dims<DIM>::ny(dim); ++y) {
                Coord<DIM> origin = Dims<DIM>::rowOrigin(y, z);
                if (!box.inBounds(origin)) {
                    continue;
                }
                std::size_t base = ((std::size_t)(tile > 0 ? z % tile : z) * Dims<DIM>::ny(dim) + y) * nx;
                for (int x = 0; x < nx; ++x) {
                    Codec<CELL>::fromRaw(&row[x], raw->data(), cells, base + x);
                }
                ret->set(Streak<DIM>(origin, nx), row.data());
            }
        }
    }

private:
    const std::vector<char> *raw;
    bool haveEdge;
    double edge;
    int tile;
};

inline std::vector<char> readFile(const char *name)
{
    FILE *f = fopen(name, "rb");
    if (!f) throw std::runtime_error(std::string("cannot open ") + name);
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<char> buf(n);
    if (n && fread(buf.data(), 1, n, f) != (size_t)n) throw std::runtime_error("short read");
    fclose(f);
    return buf;
}

template<typename CELL, typename SIM>
int runSim(const char *model, int nx, int ny, int nz, unsigned steps, const char *in, const char *out,
           bool haveEdge, double edge, const char *simName, int threads, int tile = 0)
{
    typedef typename APITraits::SelectTopology<CELL>::Value Topology;
    const int DIM = Topology::DIM;
    Coord<DIM> dim = Dims<DIM>::make(nx, ny, nz);
    std::size_t cells = (std::size_t)dim.prod();

    std::vector<char> raw = readFile(in);
    std::size_t fileCells = tile > 0 ? (std::size_t)nx * ny * tile : cells;
    if (raw.size() != fileCells * Codec<CELL>::BYTES) {
        fprintf(stderr, "input size %zu != %zu cells x %d bytes\n", raw.size(), fileCells, (int)Codec<CELL>::BYTES);
        return 2;
    }

    SIM sim(new RawInitializer<CELL>(dim, steps, &raw, haveEdge, edge, tile));
    auto t0 = std::chrono::steady_clock::now();
    sim.run();
    auto t1 = std::chrono::steady_clock::now();
    double wall = std::chrono::duration<double>(t1 - t0).count();
    double compute = sim.gatherStatistics()[0].template interval<TimeCompute>();

    const GridBase<CELL, DIM> *grid = sim.getGrid();
    const bool wantOutput = strcmp(out, "-") != 0;
    std::vector<char> res(wantOutput ? cells * Codec<CELL>::BYTES : 0);
    std::vector<CELL> row(nx);
    for (int z = 0; wantOutput && z < Dims<DIM>::nz(dim); ++z) {
        for (int y = 0; y < Dims<DIM>::ny(dim); ++y) {
            grid->get(Streak<DIM>(Dims<DIM>::rowOrigin(y, z), nx), row.data());
            std::size_t base = ((std::size_t)z * Dims<DIM>::ny(dim) + y) * nx;
            for (int x = 0; x < nx; ++x) {
                Codec<CELL>::toRaw(row[x], res.data(), cells, base + x);
            }
        }
    }
    if (wantOutput) {
        FILE *f = fopen(out, "wb");
        if (!f || fwrite(res.data(), 1, res.size(), f) != res.size()) {
            fprintf(stderr, "cannot write %s\n", out);
            return 3;
        }
        fclose(f);
    }

    unsigned nano = APITraits::SelectNanoSteps<CELL>::VALUE;
    double updates = 1.0 * steps * nano * cells;
    printf("{\"model\": \"%s\", \"simulator\": \"%s\", \"threads\": %d, \"dims\": [%d, %d, %d], \"steps\": %u, "
           "\"time_compute_s\": %.6f, \"wall_run_s\": %.6f, \"glups_compute\": %.6f, \"glups_wall\": %.6f}\n",
           model, simName, threads, nx, ny, nz, steps, compute, wall,
           1e-9 * updates / compute, 1e-9 * updates / wall);
    return 0;
}

template<typename CELL>
int runModel(const char *model, int argc, char **argv)
{
    if (argc < 8) {
        fprintf(stderr, "usage: %s <model> nx ny nz steps in.raw out.raw [--omp] [--edge v]\n", argv[0]);
        return 1;
    }
    int nx = atoi(argv[2]), ny = atoi(argv[3]), nz = atoi(argv[4]);
    unsigned steps = (unsigned)atoi(argv[5]);
    bool omp = false, haveEdge = false;
    double edge = 0;
    int tile = 0;
    for (int i = 8; i < argc; ++i) {
        if (!strcmp(argv[i], "--omp")) omp = true;
        if (!strcmp(argv[i], "--tile") && i + 1 < argc) tile = atoi(argv[++i]);
        if (!strcmp(argv[i], "--edge") && i + 1 < argc) { haveEdge = true; edge = atof(argv[++i]); }
    }
#ifdef _OPENMP
    if (omp) {
        return runSim<CELL, OpenMPSimulator<CELL> >(model, nx, ny, nz, steps, argv[6], argv[7], haveEdge, edge,
                                                    "OpenMPSimulator", omp_get_max_threads(), tile);
    }
#else
    if (omp) {
        fprintf(stderr, "built without OpenMP\n");
        return 1;
    }
#endif
    return runSim<CELL, SerialSimulator<CELL> >(model, nx, ny, nz, steps, argv[6], argv[7], haveEdge, edge,
                                                "SerialSimulator", 1, tile);
}

template<typename T>
inline T rawGet(const char *raw, std::size_t cells, int memberByteOffset, std::size_t idx)
{
    T v;
    memcpy(&v, raw + cells * memberByteOffset + idx * sizeof(T), sizeof(T));
    return v;
}

template<typename T>
inline void rawPut(char *raw, std::size_t cells, int memberByteOffset, std::size_t idx, T v)
{
    memcpy(raw + cells * memberByteOffset + idx * sizeof(T), &v, sizeof(T));
}

}

#endif
