/* TEST INFRASTRUCTURE — see oracle.h. Plain-C restatement of the reference's update path:
 *
 *   SerialSimulator::nanoStep            parallelization/serialsimulator.h:132-139
 *     one full sweep over the simulation area, then swap of the two grids;
 *   UpdateFunctor / FixedNeighborhoodUpdateFunctor
 *                                        storage/updatefunctor.h:96-139,
 *                                        storage/fixedneighborhoodupdatefunctor.h:126-256
 *     rows (streaks) in z-major, y, x order; neighbours outside a Cube read the constant edge
 *     cell kept in a padding ring of width Stencil::RADIUS (storage/soagrid.h:578-584), neighbours
 *     outside a Torus wrap (geometry/topologies.h:185-199);
 *   VanillaUpdateFunctor + CoordMap     storage/vanillaupdatefunctor.h:12-36, storage/coordmap.h:34-43
 *     same semantics per cell for the run-time-coordinate models (Game of Life).
 *
 * Both grids start from the initial state and the edge ring is never updated
 * (serialsimulator.h:54-57). Here the ring is physically present
 * (width 1) and is refilled before every sweep: with the edge constant (Cube) or with the
 * periodic image (Torus). OpenMP over planes changes nothing in the results.
 */
#include "oracle.h"

#include <stdlib.h>
#include <string.h>

static size_t pidx(int px, int py, int x, int y, int z)
{
    return ((size_t)(z + 1) * py + (y + 1)) * px + (x + 1);
}

/* fill the width-1 ring of a padded [nz+2][ny+2][nx+2] array of elem-byte items */
static void fill_ring(char *g, int elem, int nx, int ny, int nz, int torus, const void *edge)
{
    int px = nx + 2, py = ny + 2;
#pragma omp parallel for
    for (int z = -1; z <= nz; ++z) {
        for (int y = -1; y <= ny; ++y) {
            int inner_row = (z >= 0 && z < nz && y >= 0 && y < ny);
            for (int x = -1; x <= nx; x += (inner_row ? nx + 1 : 1)) {
                char *dst = g + pidx(px, py, x, y, z) * elem;
                if (torus) {
                    int sx = (x + nx) % nx, sy = (y + ny) % ny, sz = (z + nz) % nz;
                    memcpy(dst, g + pidx(px, py, sx, sy, sz) * elem, elem);
                } else {
                    memcpy(dst, edge, elem);
                }
            }
        }
    }
}

static void pad_from_dense(char *g, int elem, int nx, int ny, int nz, const char *in)
{
    int px = nx + 2, py = ny + 2;
#pragma omp parallel for
    for (int z = 0; z < nz; ++z)
        for (int y = 0; y < ny; ++y)
            memcpy(g + pidx(px, py, 0, y, z) * elem, in + ((size_t)z * ny + y) * nx * elem, (size_t)nx * elem);
}

static void dense_from_pad(const char *g, int elem, int nx, int ny, int nz, char *out)
{
    int px = nx + 2, py = ny + 2;
#pragma omp parallel for
    for (int z = 0; z < nz; ++z)
        for (int y = 0; y < ny; ++y)
            memcpy(out + ((size_t)z * ny + y) * nx * elem, g + pidx(px, py, 0, y, z) * elem, (size_t)nx * elem);
}

/* ---------------------------------------------------------------- Jacobi */

/* oracle/models/jacobi.h; arithmetic of src/examples/jacobi3d/main.cpp:33-38 (6-point) and
 * src/testbed/performancetests/main.cpp:1124-1130 (7-point order). */
int oracle_jacobi(int kind, int torus, int nx, int ny, int nz, int steps, double edge,
                  const double *in, double *out)
{
    if (kind != 6 && kind != 7 && kind != 27) return -1;
    int px = nx + 2, py = ny + 2;
    size_t n = (size_t)px * py * (nz + 2);
    double *a = malloc(n * sizeof(double)), *b = malloc(n * sizeof(double));
    if (!a || !b) return -2;
    pad_from_dense((char *)a, 8, nx, ny, nz, (const char *)in);
    memcpy(b, a, n * sizeof(double));
    const long sy = px, sz = (long)px * py;

    for (int t = 0; t < steps; ++t) {
        fill_ring((char *)a, 8, nx, ny, nz, torus, &edge);
#pragma omp parallel for
        for (int z = 0; z < nz; ++z) {
            for (int y = 0; y < ny; ++y) {
                const double *s = a + pidx(px, py, 0, y, z);
                double *d = b + pidx(px, py, 0, y, z);
                if (kind == 6) {
                    for (int x = 0; x < nx; ++x)
                        d[x] = (s[x - sz] + s[x - sy] + s[x - 1] + s[x + 1] + s[x + sy] + s[x + sz]) * (1.0 / 6.0);
                } else if (kind == 7) {
                    for (int x = 0; x < nx; ++x)
                        d[x] = (s[x - sz] + s[x - sy] + s[x - 1] + s[x] + s[x + 1] + s[x + sy] + s[x + sz]) * (1.0 / 7.0);
                } else {
#define ROW(o) ((s[x - 1 + (o)] + s[x + (o)]) + s[x + 1 + (o)])
#define PLANE(o) ((ROW((o) - sy) + ROW(o)) + ROW((o) + sy))
                    for (int x = 0; x < nx; ++x)
                        d[x] = ((PLANE(-sz) + PLANE(0)) + PLANE(sz)) * (1.0 / 27.0);
#undef ROW
#undef PLANE
                }
            }
        }
        double *tmp = a; a = b; b = tmp;
    }
    dense_from_pad((const char *)a, 8, nx, ny, nz, (char *)out);
    free(a);
    free(b);
    return 0;
}

/* ---------------------------------------------------------------- Game of Life */

/* oracle/models/conway.h; rule of src/examples/gameoflife/main.cpp:36-59. */
int oracle_gol(int torus, int nx, int ny, int steps, int edge_alive, const uint8_t *in, uint8_t *out)
{
    int px = nx + 2, py = ny + 2;
    size_t n = (size_t)px * py * 3;
    uint8_t *a = calloc(n, 1), *b = calloc(n, 1);
    if (!a || !b) return -2;
    uint8_t edge = edge_alive ? 1 : 0;
    /* a 2-D grid is handled as nz = 1; the z ring is never read */
    pad_from_dense((char *)a, 1, nx, ny, 1, (const char *)in);
    memcpy(b, a, n);

    for (int t = 0; t < steps; ++t) {
        /* only the y/x ring of plane z = 0 matters */
        for (int y = -1; y <= ny; ++y) {
            int inner = (y >= 0 && y < ny);
            for (int x = -1; x <= nx; x += (inner ? nx + 1 : 1)) {
                uint8_t v = edge;
                if (torus) v = a[pidx(px, py, (x + nx) % nx, (y + ny) % ny, 0)];
                a[pidx(px, py, x, y, 0)] = v;
            }
        }
#pragma omp parallel for
        for (int y = 0; y < ny; ++y) {
            const uint8_t *s = a + pidx(px, py, 0, y, 0);
            uint8_t *d = b + pidx(px, py, 0, y, 0);
            for (int x = 0; x < nx; ++x) {
                int living = 0;
                for (int dy = -1; dy < 2; ++dy)
                    for (int dx = -1; dx < 2; ++dx)
                        living += s[x + dx + (long)dy * px];
                int self = s[x];
                living -= self;
                d[x] = self ? ((2 <= living) && (living <= 3)) : (living == 3);
            }
        }
        uint8_t *tmp = a; a = b; b = tmp;
    }
    dense_from_pad((const char *)a, 1, nx, ny, 1, (char *)out);
    free(a);
    free(b);
    return 0;
}

/* ---------------------------------------------------------------- LBM D3Q19 */

enum { C, N, E, W, S, T, B, NW, SW, NE, SE, TW, BW, TE, BE, TN, BN, TS, BS, DENSITY, VELX, VELY, VELZ, STATE, LBM_MEMBERS };
enum { LIQUID, WEST_NOSLIP, EAST_NOSLIP, TOP, BOTTOM, NORTH_ACC, SOUTH_NOSLIP };

/* oracle/models/lbm.h; src/examples/latticeboltzmann/main.cpp:62-229. */
int oracle_lbm(int nx, int ny, int nz, int steps, const void *in_raw, void *out_raw)
{
    int px = nx + 2, py = ny + 2;
    size_t n = (size_t)px * py * (nz + 2);
    size_t cells = (size_t)nx * ny * nz;
    float *a[LBM_MEMBERS], *b[LBM_MEMBERS];
    for (int m = 0; m < LBM_MEMBERS; ++m) {
        a[m] = malloc(n * 4);
        b[m] = malloc(n * 4);
        if (!a[m] || !b[m]) return -2;
        pad_from_dense((char *)a[m], 4, nx, ny, nz, (const char *)in_raw + cells * 4 * m);
        /* edge cell = LBMCellF(): C = 1, density = 1, rest 0, state LIQUID */
        float e = (m == C || m == DENSITY) ? 1.0f : 0.0f;
        int32_t ei = 0;
        fill_ring((char *)a[m], 4, nx, ny, nz, 0, m == STATE ? (void *)&ei : (void *)&e);
        memcpy(b[m], a[m], n * 4);
    }
    const long sy = px, sz = (long)px * py;

    const float omega     = (float)(1.0 / 1.7);
    const float omega_trm = 1.0f - omega;
    const float omega_w0  = (float)(3.0 * 1.0 / 3.0)  * omega;
    const float omega_w1  = (float)(3.0 * 1.0 / 18.0) * omega;
    const float omega_w2  = (float)(3.0 * 1.0 / 36.0) * omega;
    const float one_third = (float)(1.0 / 3.0);

#define GET_COMP(X, Y, Z, COMP) a[COMP][i + (X) + (Y) * sy + (Z) * sz]
#define SQR(X) ((X) * (X))
    for (int t = 0; t < steps; ++t) {
#pragma omp parallel for
        for (int z = 0; z < nz; ++z) {
            for (int y = 0; y < ny; ++y) {
                for (int x = 0; x < nx; ++x) {
                    const long i = (long)pidx(px, py, x, y, z);
                    int32_t s;
                    memcpy(&s, &a[STATE][i], 4);
                    if (s != LIQUID) {
                        for (int m = 0; m < LBM_MEMBERS; ++m) b[m][i] = a[m][i];
                        switch (s) {
                        case WEST_NOSLIP:
                            b[E][i]  = GET_COMP(1, 0,  0, W);
                            b[NE][i] = GET_COMP(1, 1,  0, SW);
                            b[SE][i] = GET_COMP(1,-1,  0, NW);
                            b[TE][i] = GET_COMP(1, 0,  1, BW);
                            b[BE][i] = GET_COMP(1, 0, -1, TW);
                            break;
                        case EAST_NOSLIP:
                            b[W][i]  = GET_COMP(-1, 0, 0, E);
                            b[NW][i] = GET_COMP(-1, 0, 1, SE);
                            b[SW][i] = GET_COMP(-1,-1, 0, NE);
                            b[TW][i] = GET_COMP(-1, 0, 1, BE);
                            b[BW][i] = GET_COMP(-1, 0,-1, TE);
                            break;
                        case TOP:
                            b[B][i]  = GET_COMP(0, 0,-1, T);
                            b[BE][i] = GET_COMP(1, 0,-1, TW);
                            b[BW][i] = GET_COMP(-1,0,-1, TE);
                            b[BN][i] = GET_COMP(0, 1,-1, TS);
                            b[BS][i] = GET_COMP(0,-1,-1, TN);
                            break;
                        case BOTTOM:
                            b[T][i]  = GET_COMP(0, 0, 1, B);
                            b[TE][i] = GET_COMP(1, 0, 1, BW);
                            b[TW][i] = GET_COMP(-1,0, 1, BE);
                            b[TN][i] = GET_COMP(0, 1, 1, BS);
                            b[TS][i] = GET_COMP(0,-1, 1, BN);
                            break;
                        case NORTH_ACC: {
                            const float w_1 = 0.01f;
                            b[S][i]  = GET_COMP(0,-1, 0, N);
                            b[SE][i] = GET_COMP(1,-1, 0, NW) + 6.0f * w_1 * 0.1f;
                            b[SW][i] = GET_COMP(-1,-1,0, NE) - 6.0f * w_1 * 0.1f;
                            b[TS][i] = GET_COMP(0,-1, 1, BN);
                            b[BS][i] = GET_COMP(0,-1,-1, TN);
                            break;
                        }
                        case SOUTH_NOSLIP:
                            b[N][i]  = GET_COMP(0, 1, 0, S);
                            b[NE][i] = GET_COMP(1, 1, 0, SW);
                            b[NW][i] = GET_COMP(-1,1, 0, SE);
                            b[TN][i] = GET_COMP(0, 1, 1, BS);
                            b[BN][i] = GET_COMP(0, 1,-1, TS);
                            break;
                        }
                        continue;
                    }

                    float velX, velY, velZ;
                    velX =
                        GET_COMP(-1, 0, 0, E)  + GET_COMP(-1,-1, 0, NE) +
                        GET_COMP(-1, 1, 0, SE) + GET_COMP(-1, 0,-1, TE) +
                        GET_COMP(-1, 0, 1, BE);
                    velY = GET_COMP(0,-1, 0, N) + GET_COMP(1,-1, 0, NW) +
                        GET_COMP(0,-1,-1, TN) + GET_COMP(0,-1, 1, BN);
                    velZ = GET_COMP(0, 0,-1, T) + GET_COMP(0, 1,-1, TS) +
                        GET_COMP(1, 0,-1, TW);

                    const float rho =
                        GET_COMP(0, 0, 0, C)  + GET_COMP(0, 1, 0, S) +
                        GET_COMP(1, 0, 0, W)  + GET_COMP(0, 0, 1, B) +
                        GET_COMP(1, 1, 0, SW) + GET_COMP(0, 1, 1, BS) +
                        GET_COMP(1, 0, 1, BW) + velX + velY + velZ;
                    velX = velX
                        - GET_COMP(1, 0, 0, W)  - GET_COMP(1,-1, 0, NW)
                        - GET_COMP(1, 1, 0, SW) - GET_COMP(1, 0,-1, TW)
                        - GET_COMP(1, 0, 1, BW);
                    velY = velY
                        + GET_COMP(-1,-1, 0, NE) - GET_COMP(0, 1, 0, S)
                        - GET_COMP(1, 1, 0, SW)  - GET_COMP(-1, 1, 0, SE)
                        - GET_COMP(0, 1,-1, TS)  - GET_COMP(0, 1, 1, BS);
                    velZ = velZ + GET_COMP(0,-1,-1, TN) + GET_COMP(-1, 0,-1, TE) - GET_COMP(0, 0, 1, B)
                        - GET_COMP(0,-1, 1, BN) - GET_COMP(0, 1, 1, BS) - GET_COMP(1, 0, 1, BW)
                        - GET_COMP(-1, 0, 1, BE);

                    b[DENSITY][i] = rho;
                    b[VELX][i] = velX;
                    b[VELY][i] = velY;
                    b[VELZ][i] = velZ;

                    const float dir_indep_trm = one_third * rho - 0.5f * (velX * velX + velY * velY + velZ * velZ);

                    b[C][i]  = omega_trm * GET_COMP(0, 0, 0, C) + omega_w0 * (dir_indep_trm);

                    b[NW][i] = omega_trm * GET_COMP( 1,-1, 0, NW) + omega_w2 * (dir_indep_trm - (velX - velY) + 1.5f * SQR(velX - velY));
                    b[SE][i] = omega_trm * GET_COMP(-1, 1, 0, SE) + omega_w2 * (dir_indep_trm + (velX - velY) + 1.5f * SQR(velX - velY));
                    b[NE][i] = omega_trm * GET_COMP(-1,-1, 0, NE) + omega_w2 * (dir_indep_trm + (velX + velY) + 1.5f * SQR(velX + velY));
                    b[SW][i] = omega_trm * GET_COMP( 1, 1, 0, SW) + omega_w2 * (dir_indep_trm - (velX + velY) + 1.5f * SQR(velX + velY));

                    b[TW][i] = omega_trm * GET_COMP( 1, 0,-1, TW) + omega_w2 * (dir_indep_trm - (velX - velZ) + 1.5f * SQR(velX - velZ));
                    b[BE][i] = omega_trm * GET_COMP(-1, 0, 1, BE) + omega_w2 * (dir_indep_trm + (velX - velZ) + 1.5f * SQR(velX - velZ));
                    b[TE][i] = omega_trm * GET_COMP(-1, 0,-1, TE) + omega_w2 * (dir_indep_trm + (velX + velZ) + 1.5f * SQR(velX + velZ));
                    b[BW][i] = omega_trm * GET_COMP( 1, 0, 1, BW) + omega_w2 * (dir_indep_trm - (velX + velZ) + 1.5f * SQR(velX + velZ));

                    b[TS][i] = omega_trm * GET_COMP(0, 1,-1, TS) + omega_w2 * (dir_indep_trm - (velY - velZ) + 1.5f * SQR(velY - velZ));
                    b[BN][i] = omega_trm * GET_COMP(0,-1, 1, BN) + omega_w2 * (dir_indep_trm + (velY - velZ) + 1.5f * SQR(velY - velZ));
                    b[TN][i] = omega_trm * GET_COMP(0,-1,-1, TN) + omega_w2 * (dir_indep_trm + (velY + velZ) + 1.5f * SQR(velY + velZ));
                    b[BS][i] = omega_trm * GET_COMP(0, 1, 1, BS) + omega_w2 * (dir_indep_trm - (velY + velZ) + 1.5f * SQR(velY + velZ));

                    b[N][i] = omega_trm * GET_COMP(0,-1, 0, N) + omega_w1 * (dir_indep_trm + velY + 1.5f * SQR(velY));
                    b[S][i] = omega_trm * GET_COMP(0, 1, 0, S) + omega_w1 * (dir_indep_trm - velY + 1.5f * SQR(velY));
                    b[E][i] = omega_trm * GET_COMP(-1, 0, 0, E) + omega_w1 * (dir_indep_trm + velX + 1.5f * SQR(velX));
                    b[W][i] = omega_trm * GET_COMP( 1, 0, 0, W) + omega_w1 * (dir_indep_trm - velX + 1.5f * SQR(velX));
                    b[T][i] = omega_trm * GET_COMP(0, 0,-1, T) + omega_w1 * (dir_indep_trm + velZ + 1.5f * SQR(velZ));
                    b[B][i] = omega_trm * GET_COMP(0, 0, 1, B) + omega_w1 * (dir_indep_trm - velZ + 1.5f * SQR(velZ));
                    b[STATE][i] = a[STATE][i];
                }
            }
        }
        for (int m = 0; m < LBM_MEMBERS; ++m) {
            float *tmp = a[m]; a[m] = b[m]; b[m] = tmp;
        }
    }
#undef GET_COMP
#undef SQR
    for (int m = 0; m < LBM_MEMBERS; ++m) {
        dense_from_pad((const char *)a[m], 4, nx, ny, nz, (char *)out_raw + cells * 4 * m);
        free(a[m]);
        free(b[m]);
    }
    return 0;
}

/* ---------------------------------------------------------------- region (de)serialisation */

/* SoAGrid::saveRegion -> LFA soa_grid::save -> save_functor (storage/soagrid.h:523-547,
 * lib/libflatarray/include/libflatarray/detail/save_functor.hpp:26-66): per member, the cells of
 * all streaks in streak order, tightly packed; the buffer is member-major with stride =
 * total cell count of the region. */
static int region_copy(int nx, int ny, int nz, int n_members, const int *member_bytes,
                       char *grid_raw, const int *streaks, int n_streaks, char *buf, int save)
{
    size_t cells = (size_t)nx * ny * nz, count = 0;
    for (int s = 0; s < n_streaks; ++s) {
        const int *k = streaks + 4 * s;
        if (k[0] < 0 || k[3] > nx || k[3] < k[0] || k[1] < 0 || k[1] >= ny || k[2] < 0 || k[2] >= nz) return -1;
        count += (size_t)(k[3] - k[0]);
    }
    size_t goff = 0, boff = 0;
    for (int m = 0; m < n_members; ++m) {
        size_t eb = (size_t)member_bytes[m], pos = 0;
        for (int s = 0; s < n_streaks; ++s) {
            const int *k = streaks + 4 * s;
            size_t len = (size_t)(k[3] - k[0]);
            char *g = grid_raw + goff + (((size_t)k[2] * ny + k[1]) * nx + k[0]) * eb;
            char *p = buf + boff + pos * eb;
            if (save) memcpy(p, g, len * eb); else memcpy(g, p, len * eb);
            pos += len;
        }
        goff += cells * eb;
        boff += count * eb;
    }
    return 0;
}

int oracle_save_region(int nx, int ny, int nz, int n_members, const int *member_bytes,
                       const void *grid_raw, const int *streaks, int n_streaks, void *buf)
{
    return region_copy(nx, ny, nz, n_members, member_bytes, (char *)grid_raw, streaks, n_streaks, (char *)buf, 1);
}

int oracle_load_region(int nx, int ny, int nz, int n_members, const int *member_bytes,
                       void *grid_raw, const int *streaks, int n_streaks, const void *buf)
{
    return region_copy(nx, ny, nz, n_members, member_bytes, (char *)grid_raw, streaks, n_streaks, (char *)buf, 0);
}

/* ------------------------------------------------------------------------------------------
 * Short-range n-body in BoxCell containers (oracle/models/nbody.h):
 *   BoxCell::update / copyOver / addContainedParticles / updateCargo   storage/boxcell.h:112-174
 *   SelectPositionChecker (origin <= pos < origin + dimension, doubles) misc/apitraits.h:1074-1088
 *   NeighborhoodIterator: the 27 cells in CoordBox order (x fastest), each cell's particles in
 *   storage order; cells outside the Cube are the empty edge container
 *                                                                       storage/neighborhooditerator.h:71-186
 *   FixedArray::operator<< throws std::out_of_range("capacity exceeded")  storage/fixedarray.h:77-83
 * NANO_STEPS = 1, so every step re-bins (nanoStep == 0) and then updates each particle of the NEW
 * container against the OLD grid.
 */
#define DEFINE_NBODY(NAME, REAL)                                                                   \
static int NAME(int nx, int ny, int nz, int cap, int steps, double dt_, double cutoff, double edge,\
                const int *org, const int32_t *cin, const REAL *pin, int32_t *cout, REAL *pout)    \
{                                                                                                  \
    size_t cells = (size_t)nx * ny * nz;                                                           \
    int32_t *c0 = malloc(cells * sizeof(int32_t)), *c1 = malloc(cells * sizeof(int32_t));          \
    REAL *p0 = malloc(cells * cap * 6 * sizeof(REAL)), *p1 = malloc(cells * cap * 6 * sizeof(REAL)); \
    if (!c0 || !c1 || !p0 || !p1) { free(c0); free(c1); free(p0); free(p1); return -5; }           \
    memcpy(c0, cin, cells * sizeof(int32_t));                                                      \
    memcpy(p0, pin, cells * cap * 6 * sizeof(REAL));                                               \
    const REAL dt = (REAL)dt_, rc = (REAL)cutoff, rc2 = rc * rc;                                   \
    int overflow = 0;                                                                              \
    for (int s = 0; s < steps; ++s) {                                                              \
        _Pragma("omp parallel for collapse(2) reduction(|:overflow)")                              \
        for (int z = 0; z < nz; ++z) {                                                             \
            for (int y = 0; y < ny; ++y) {                                                         \
                for (int x = 0; x < nx; ++x) {                                                     \
                    size_t self = ((size_t)z * ny + y) * nx + x;                                   \
                    double o[3] = {(x + org[0]) * edge, (y + org[1]) * edge, (z + org[2]) * edge}; \
                    double q[3] = {o[0] + edge, o[1] + edge, o[2] + edge};                         \
                    REAL *mine = p1 + self * cap * 6;                                              \
                    int n = 0;                                                                     \
                    /* copyOver at nanoStep 0: clear, then addContainedParticles over the hood */  \
                    for (int k = 0; k < 27; ++k) {                                                 \
                        int ax = x + k % 3 - 1, ay = y + (k / 3) % 3 - 1, az = z + k / 9 - 1;      \
                        if (ax < 0 || ax >= nx || ay < 0 || ay >= ny || az < 0 || az >= nz) continue; \
                        size_t nb = ((size_t)az * ny + ay) * nx + ax;                              \
                        for (int p = 0; p < c0[nb]; ++p) {                                         \
                            const REAL *src = p0 + (nb * cap + p) * 6;                             \
                            double px = src[0], py = src[1], pz = src[2];                          \
                            if (o[0] <= px && o[1] <= py && o[2] <= pz && px < q[0] && py < q[1] && pz < q[2]) { \
                                if (n >= cap) { overflow = 1; continue; }                          \
                                memcpy(mine + n * 6, src, 6 * sizeof(REAL));                       \
                                ++n;                                                               \
                            }                                                                      \
                        }                                                                          \
                    }                                                                              \
                    c1[self] = n;                                                                  \
                    /* updateCargo: Particle::update against the OLD neighbourhood */              \
                    for (int i = 0; i < n; ++i) {                                                  \
                        REAL *t = mine + i * 6;                                                    \
                        for (int k = 0; k < 27; ++k) {                                             \
                            int ax = x + k % 3 - 1, ay = y + (k / 3) % 3 - 1, az = z + k / 9 - 1;  \
                            if (ax < 0 || ax >= nx || ay < 0 || ay >= ny || az < 0 || az >= nz) continue; \
                            size_t nb = ((size_t)az * ny + ay) * nx + ax;                          \
                            for (int p = 0; p < c0[nb]; ++p) {                                     \
                                const REAL *src = p0 + (nb * cap + p) * 6;                         \
                                REAL d0 = t[0] - src[0], d1 = t[1] - src[1], d2 = t[2] - src[2];   \
                                REAL r2 = (d0 * d0 + d1 * d1) + d2 * d2;                           \
                                if (r2 == 0 || r2 >= rc2) continue;                                \
                                REAL inv = (REAL)1 / r2;                                           \
                                REAL s6 = inv * inv * inv;                                         \
                                REAL f = ((REAL)24 * inv) * s6 * ((REAL)2 * s6 - (REAL)1);         \
                                t[3] += (d0 * f) * dt;                                             \
                                t[4] += (d1 * f) * dt;                                             \
                                t[5] += (d2 * f) * dt;                                             \
                            }                                                                      \
                        }                                                                          \
                        t[0] += t[3] * dt;                                                         \
                        t[1] += t[4] * dt;                                                         \
                        t[2] += t[5] * dt;                                                         \
                    }                                                                              \
                }                                                                                  \
            }                                                                                      \
        }                                                                                          \
        if (overflow) break;                                                                       \
        int32_t *tc = c0; c0 = c1; c1 = tc;                                                        \
        REAL *tp = p0; p0 = p1; p1 = tp;                                                           \
    }                                                                                              \
    if (!overflow) {                                                                               \
        memcpy(cout, c0, cells * sizeof(int32_t));                                                 \
        memset(pout, 0, cells * cap * 6 * sizeof(REAL));                                           \
        for (size_t c = 0; c < cells; ++c)                                                         \
            memcpy(pout + c * cap * 6, p0 + c * cap * 6, (size_t)c0[c] * 6 * sizeof(REAL));        \
    }                                                                                              \
    free(c0); free(c1); free(p0); free(p1);                                                        \
    return overflow ? -3 : 0;                                                                      \
}

DEFINE_NBODY(nbody_f32, float)
DEFINE_NBODY(nbody_f64, double)

int oracle_nbody_at(int real_bytes, int nx, int ny, int nz, int cap, int steps, double dt, double cutoff, double edge,
                    const int origin[3], const int32_t *counts_in, const void *parts_in, int32_t *counts_out, void *parts_out)
{
    if (real_bytes == 4)
        return nbody_f32(nx, ny, nz, cap, steps, dt, cutoff, edge, origin, counts_in, (const float *)parts_in, counts_out, (float *)parts_out);
    if (real_bytes == 8)
        return nbody_f64(nx, ny, nz, cap, steps, dt, cutoff, edge, origin, counts_in, (const double *)parts_in, counts_out, (double *)parts_out);
    return -1;
}

int oracle_nbody(int real_bytes, int nx, int ny, int nz, int cap, int steps, double dt, double cutoff, double edge,
                 const int32_t *counts_in, const void *parts_in, int32_t *counts_out, void *parts_out)
{
    const int origin[3] = {0, 0, 0};
    return oracle_nbody_at(real_bytes, nx, ny, nz, cap, steps, dt, cutoff, edge, origin, counts_in, parts_in, counts_out, parts_out);
}

/* ---- ContainerCell: ID-keyed mesh elements ------------------------------------------------------ */

/* ContainerCell::operator[](id), storage/containercell.h:107-121: upper_bound over the ascending ids, then the
 * element before it */
static int container_find(const int32_t *ids, int n, int32_t id)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = lo + (hi - lo) / 2;
        if (id < ids[mid]) hi = mid;
        else lo = mid + 1;
    }
    if (lo == 0) return -1;
    return ids[lo - 1] == id ? lo - 1 : -1;
}

int oracle_container(int n_dims, int torus, int nx, int ny, int nz, int cap, int maxnb, int steps,
                     const int32_t *counts, const int32_t *ids, const double *values, const double *influx,
                     const int32_t *nb_counts, const int32_t *nb_ids,
                     const int32_t *edge_count, const int32_t *edge_ids, const double *edge_values,
                     double *values_out, int32_t *missing_id)
{
    if ((n_dims != 2 && n_dims != 3) || (n_dims == 2 && nz != 1) || nx < 1 || ny < 1 || nz < 1 || cap < 1 || maxnb < 1) return -1;
    size_t cells = (size_t)nx * ny * nz;
    double *cur = (double *)malloc(cells * cap * sizeof(double));
    double *next = (double *)malloc(cells * cap * sizeof(double));
    if (!cur || !next) {
        free(cur);
        free(next);
        return -1;
    }
    memcpy(cur, values, cells * cap * sizeof(double));
    int n_edge = edge_count ? edge_count[0] : 0;
    int rc = 0;
    int zlo = n_dims == 3 ? -1 : 0, zhi = n_dims == 3 ? 1 : 0;
    for (int t = 0; t < steps && rc == 0; ++t) {
        /* copyOver: *this = oldSelf (containercell.h:185-188); elements keep everything but the temperature */
        memcpy(next, cur, cells * cap * sizeof(double));
        for (int z = 0; z < nz && rc == 0; ++z) {
            for (int y = 0; y < ny && rc == 0; ++y) {
                for (int x = 0; x < nx && rc == 0; ++x) {
                    size_t c = ((size_t)z * ny + y) * nx + x;
                    /* updateCargo (containercell.h:195-200) */
                    for (int s = 0; s < counts[c] && rc == 0; ++s) {
                        size_t slot = c * cap + s;
                        double temperature = 0;
                        for (int j = 0; j < nb_counts[slot]; ++j) {
                            int32_t id = nb_ids[slot * maxnb + j];
                            /* NeighborhoodAdapter::operator[] (neighborhoodadapter.h:45-65): own container first */
                            const double *hit = 0;
                            int pos = container_find(ids + c * cap, counts[c], id);
                            if (pos >= 0) hit = cur + c * cap + pos;
                            for (int dz = zlo; dz <= zhi && !hit; ++dz) {
                                for (int dy = -1; dy <= 1 && !hit; ++dy) {
                                    for (int dx = -1; dx <= 1 && !hit; ++dx) {
                                        if (!dx && !dy && !dz) continue;
                                        int p[3] = {x + dx, y + dy, z + dz};
                                        int d[3] = {nx, ny, nz};
                                        int outside = 0;
                                        for (int a = 0; a < 3; ++a) {
                                            if (p[a] < 0 || p[a] >= d[a]) {
                                                if (torus) p[a] = (p[a] + d[a]) % d[a];
                                                else outside = 1;
                                            }
                                        }
                                        if (outside) {
                                            /* Cube: every coordinate outside is the edge cell (geometry/topologies.h:185-199) */
                                            pos = container_find(edge_ids, n_edge, id);
                                            if (pos >= 0) hit = edge_values + pos;
                                        } else {
                                            size_t o = ((size_t)p[2] * ny + p[1]) * nx + p[0];
                                            pos = container_find(ids + o * cap, counts[o], id);
                                            if (pos >= 0) hit = cur + o * cap + pos;
                                        }
                                    }
                                }
                            }
                            if (!hit) {
                                if (missing_id) *missing_id = id;
                                rc = -2;
                                break;
                            }
                            temperature += *hit;
                        }
                        if (rc) break;
                        /* src/examples/voronoi/main.cpp:53: the size_t divisor converts to double */
                        next[slot] = influx[slot] + temperature / (double)(size_t)nb_counts[slot];
                    }
                }
            }
        }
        double *tmp = cur;
        cur = next;
        next = tmp;
    }
    if (rc == 0) {
        for (size_t c = 0; c < cells; ++c)
            for (int s = 0; s < cap; ++s) values_out[c * cap + s] = s < counts[c] ? cur[c * cap + s] : 0.0;
    }
    free(cur);
    free(next);
    return rc;
}
