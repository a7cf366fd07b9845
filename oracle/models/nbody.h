/* TEST INFRASTRUCTURE — short-range n-body model written against the UNCHANGED LibGeoDecomp
 * plugin API: particles live in BoxCell<FixedArray<Particle, N>> containers
 * (storage/boxcell.h:21-178, storage/fixedarray.h:21-140), one container per grid cell of edge
 * `cellEdge` >= cutoff; BoxCell re-bins them from the 27-cell Moore neighbourhood at nano step 0
 * (boxcell.h:123-138,164-174, position checker misc/apitraits.h:1074-1088) and then calls
 * Particle::update(hood, nanoStep) for each, the hood iterating over all particles of the 27
 * cells in CoordBox order (storage/neighborhooditerator.h:71-186).
 *
 * The particle follows src/examples/bouncingspheres/main.cpp:92-128 (accumulate force * dt into
 * the velocity neighbour by neighbour, then pos += vel * dt) with a Lennard-Jones 12-6 force
 * truncated at the cutoff instead of the soft-sphere repulsion. The same source is the model the
 * CUDA kernel (libgeodecomp_b200/csrc/nbody.cu) restates; built -ffp-contract=off.
 *
 * Raw file format of the driver below (one rank, whole grid):
 *   int32  counts[nz][ny][nx]
 *   REAL   particles[nz][ny][nx][N][6]      pos x,y,z, vel x,y,z; unused slots are ignored
 * Cell (x, y, z) has origin (x, y, z) * cellEdge and dimension cellEdge^3 (double arithmetic).
 */
#ifndef B200GEO_ORACLE_MODELS_NBODY_H
#define B200GEO_ORACLE_MODELS_NBODY_H

#include <libgeodecomp/geometry/floatcoord.h>
#include <libgeodecomp/misc/apitraits.h>
#include <libgeodecomp/storage/boxcell.h>
#include <libgeodecomp/storage/fixedarray.h>

namespace b200models {

using namespace LibGeoDecomp;

struct NBodyParams {
    static double& dt() { static double v = 0.005; return v; }
    static double& cutoff() { static double v = 2.5; return v; }
    static double& cellEdge() { static double v = 2.5; return v; }
};

const int NBODY_CAPACITY = 32;

template<typename REAL>
class LJParticle
{
public:
    class API :
        public APITraits::HasCubeTopology<3>,
        public APITraits::HasStencil<Stencils::Moore<3, 1> >
    {};

    LJParticle()
    {
        for (int k = 0; k < 3; ++k) {
            pos[k] = 0;
            vel[k] = 0;
        }
    }

    FloatCoord<3> getPos() const
    {
        return FloatCoord<3>(pos[0], pos[1], pos[2]);
    }

    template<typename HOOD>
    void update(const HOOD& hood, const int /* nanoStep */)
    {
        const REAL dt = (REAL)NBodyParams::dt();
        const REAL rc = (REAL)NBodyParams::cutoff();
        const REAL rc2 = rc * rc;

        for (typename HOOD::Iterator i = hood.begin(); i != hood.end(); ++i) {
            const LJParticle& other = *i;
            REAL d0 = pos[0] - other.pos[0];
            REAL d1 = pos[1] - other.pos[1];
            REAL d2 = pos[2] - other.pos[2];
            REAL r2 = (d0 * d0 + d1 * d1) + d2 * d2;
            if ((r2 == 0) || (r2 >= rc2)) {
                continue;
            }
            REAL inv = (REAL)1 / r2;
            REAL s6 = inv * inv * inv;
            REAL f = ((REAL)24 * inv) * s6 * ((REAL)2 * s6 - (REAL)1);
            vel[0] += (d0 * f) * dt;
            vel[1] += (d1 * f) * dt;
            vel[2] += (d2 * f) * dt;
        }

        pos[0] += vel[0] * dt;
        pos[1] += vel[1] * dt;
        pos[2] += vel[2] * dt;
    }

    REAL pos[3];
    REAL vel[3];
};

typedef BoxCell<FixedArray<LJParticle<float>, NBODY_CAPACITY> > NBodyCellF;
typedef BoxCell<FixedArray<LJParticle<double>, NBODY_CAPACITY> > NBodyCellD;
typedef NBodyCellF NBodyCell;

}

#ifdef B200GEO_ORACLE_REF_DRIVER_H

namespace b200models {

using namespace refdriver;

template<typename REAL>
class NBodyInitializer : public SimpleInitializer<BoxCell<FixedArray<LJParticle<REAL>, NBODY_CAPACITY> > >
{
public:
    typedef BoxCell<FixedArray<LJParticle<REAL>, NBODY_CAPACITY> > Cell;

    NBodyInitializer(const Coord<3>& dim, unsigned steps, const std::vector<char> *raw) :
        SimpleInitializer<Cell>(dim, steps), raw(raw)
    {}

    virtual void grid(GridBase<Cell, 3> *ret)
    {
        Coord<3> dim = this->gridDimensions();
        CoordBox<3> box = ret->boundingBox();
        std::size_t cells = (std::size_t)dim.prod();
        const int *counts = (const int*)raw->data();
        const REAL *parts = (const REAL*)(raw->data() + cells * sizeof(int));
        double edge = NBodyParams::cellEdge();

        for (CoordBox<3>::Iterator i = box.begin(); i != box.end(); ++i) {
            Coord<3> c = *i;
            if (!CoordBox<3>(Coord<3>(), dim).inBounds(c)) {
                continue;
            }
            std::size_t idx = ((std::size_t)c.z() * dim.y() + c.y()) * dim.x() + c.x();
            Cell cell(FloatCoord<3>(c.x() * edge, c.y() * edge, c.z() * edge), FloatCoord<3>(edge, edge, edge));
            for (int p = 0; p < counts[idx]; ++p) {
                LJParticle<REAL> particle;
                const REAL *src = parts + (idx * NBODY_CAPACITY + p) * 6;
                for (int k = 0; k < 3; ++k) {
                    particle.pos[k] = src[k];
                    particle.vel[k] = src[3 + k];
                }
                cell << particle;
            }
            ret->set(c, cell);
        }
    }

private:
    const std::vector<char> *raw;
};

template<typename REAL, typename SIM>
int nbodyRun(int nx, int ny, int nz, unsigned steps, const char *in, const char *out, const char *simName, int threads)
{
    typedef BoxCell<FixedArray<LJParticle<REAL>, NBODY_CAPACITY> > Cell;
    Coord<3> dim(nx, ny, nz);
    std::size_t cells = (std::size_t)dim.prod();
    std::vector<char> raw = readFile(in);
    std::size_t want = cells * sizeof(int) + cells * NBODY_CAPACITY * 6 * sizeof(REAL);
    if (raw.size() != want) {
        fprintf(stderr, "input size %zu != %zu\n", raw.size(), want);
        return 2;
    }
    std::size_t particles = 0;
    for (std::size_t i = 0; i < cells; ++i) particles += ((const int*)raw.data())[i];

    SIM sim(new NBodyInitializer<REAL>(dim, steps, &raw));
    auto t0 = std::chrono::steady_clock::now();
    sim.run();
    auto t1 = std::chrono::steady_clock::now();
    double wall = std::chrono::duration<double>(t1 - t0).count();
    double compute = sim.gatherStatistics()[0].template interval<TimeCompute>();

    const GridBase<Cell, 3> *grid = sim.getGrid();
    std::vector<char> res(raw.size(), 0);
    int *counts = (int*)res.data();
    REAL *parts = (REAL*)(res.data() + cells * sizeof(int));
    std::size_t left = 0;
    for (int z = 0; z < nz; ++z) {
        for (int y = 0; y < ny; ++y) {
            for (int x = 0; x < nx; ++x) {
                Cell cell = grid->get(Coord<3>(x, y, z));
                std::size_t idx = ((std::size_t)z * ny + y) * nx + x;
                counts[idx] = (int)cell.size();
                left += cell.size();
                for (std::size_t p = 0; p < cell.size(); ++p) {
                    REAL *dst = parts + (idx * NBODY_CAPACITY + p) * 6;
                    for (int k = 0; k < 3; ++k) {
                        dst[k] = cell[p].pos[k];
                        dst[3 + k] = cell[p].vel[k];
                    }
                }
            }
        }
    }
    if (strcmp(out, "-") != 0) {
        FILE *f = fopen(out, "wb");
        if (!f || fwrite(res.data(), 1, res.size(), f) != res.size()) {
            fprintf(stderr, "cannot write %s\n", out);
            return 3;
        }
        fclose(f);
    }
    double updates = 1.0 * steps * particles;
    printf("{\"model\": \"nbody\", \"real_bytes\": %d, \"simulator\": \"%s\", \"threads\": %d, \"dims\": [%d, %d, %d], "
           "\"steps\": %u, \"particles\": %zu, \"particles_left\": %zu, \"time_compute_s\": %.6f, \"wall_run_s\": %.6f, "
           "\"gpups_compute\": %.6f, \"glups_compute\": %.6f}\n",
           (int)sizeof(REAL), simName, threads, nx, ny, nz, steps, particles, left, compute, wall,
           1e-9 * updates / compute, 1e-9 * steps * cells / compute);
    return 0;
}

/* usage: lgd_ref_nbody nbody nx ny nz steps in.raw out.raw [--omp] [--double] [--dt v] [--cutoff v] [--edge v] */
inline int nbodyMain(int argc, char **argv)
{
    if (argc < 8) {
        fprintf(stderr, "usage: %s nbody nx ny nz steps in.raw out.raw [--omp] [--double] [--dt v] [--cutoff v] [--edge v]\n", argv[0]);
        return 1;
    }
    int nx = atoi(argv[2]), ny = atoi(argv[3]), nz = atoi(argv[4]);
    unsigned steps = (unsigned)atoi(argv[5]);
    bool omp = false, dbl = false;
    for (int i = 8; i < argc; ++i) {
        if (!strcmp(argv[i], "--omp")) omp = true;
        if (!strcmp(argv[i], "--double")) dbl = true;
        if (!strcmp(argv[i], "--dt") && i + 1 < argc) NBodyParams::dt() = atof(argv[++i]);
        if (!strcmp(argv[i], "--cutoff") && i + 1 < argc) NBodyParams::cutoff() = atof(argv[++i]);
        if (!strcmp(argv[i], "--edge") && i + 1 < argc) NBodyParams::cellEdge() = atof(argv[++i]);
    }
    try {
#ifdef _OPENMP
        if (omp) {
            int t = omp_get_max_threads();
            return dbl ? nbodyRun<double, OpenMPSimulator<NBodyCellD> >(nx, ny, nz, steps, argv[6], argv[7], "OpenMPSimulator", t)
                       : nbodyRun<float, OpenMPSimulator<NBodyCellF> >(nx, ny, nz, steps, argv[6], argv[7], "OpenMPSimulator", t);
        }
#endif
        return dbl ? nbodyRun<double, SerialSimulator<NBodyCellD> >(nx, ny, nz, steps, argv[6], argv[7], "SerialSimulator", 1)
                   : nbodyRun<float, SerialSimulator<NBodyCellF> >(nx, ny, nz, steps, argv[6], argv[7], "SerialSimulator", 1);
    } catch (const std::exception& e) {
        fprintf(stderr, "error: %s\n", e.what());
        return 4;
    }
}

}

#endif

#endif
