// Short-range n-body in BoxCell containers (B200GEO_KERNEL_NBODY): device-resident cell-list
// grid, re-binning and particle update.
//
// Replaces, for the bound particle model (oracle/models/nbody.h: Lennard-Jones 12-6 truncated at the
// cutoff, velocity += force * dt neighbour by neighbour, then pos += vel * dt), the reference's
//   BoxCell::update / copyOver / addContainedParticles / updateCargo   storage/boxcell.h:112-174
//   SelectPositionChecker (origin <= pos < origin + dimension, doubles) misc/apitraits.h:1074-1088
//   NeighborhoodIterator (27 cells in CoordBox order, x fastest; particles in storage order)
//                                                                       storage/neighborhooditerator.h:71-186
//   FixedArray<Particle, N> (capacity N, operator<< throws std::out_of_range when full)
//                                                                       storage/fixedarray.h:21-83
// Same expression trees, same neighbour order, compiled -fmad=false against a -ffp-contract=off
// oracle: positions and velocities are bit-identical to the reference's SerialSimulator.
//
// Layout in HBM (per buffer; two buffers, swapped per sweep like the stencil grids):
//   counts    int32 [nz+2][ny+2][nx+2]            one ring of ghost containers: EDGE = the empty
//   particles REAL  [nz+2][ny+2][nx+2][6][cap]     edge container of a Cube, PEER = a slab neighbour's
// i.e. SoA inside a container (x[cap], y[cap], z[cap], vx[cap], vy[cap], vz[cap]): a warp reading
// one component of one container touches one or two 128-byte lines, and a z-slab's ghost plane is
// one contiguous block per array (halo exchange in place, no pack kernel).
//
// Roofline: FP32/FP64 pipe, not HBM (about 27 * <n> candidate pairs per particle at ~35 flop
// versus 24 + 24 bytes per particle of traffic) and no tensor cores (no dense contraction).
//
// Two kernels per sweep:
//   rebin_kernel   one warp per container: scans the 27 old containers in order, keeps the
//                  particles whose position lies in this container's box (ordered compaction by
//                  ballot/popc), writes the new container;
//   force_kernel   one CTA per run of G x-adjacent containers: stages the positions of the
//                  (G + 2) x 3 x 3 surrounding OLD containers in shared memory, packs the run's new
//                  particles densely onto threads (so lanes are not wasted on empty slots), and
//                  each thread walks its particle's 27 containers in reference order.
#include "grid.h"

#include <cstring>
#include <new>

struct b200geo_boxgrid {
    b200geo_boxgrid_desc desc;
    int device;
    int d[3];
    int cap, real;
    int64_t pcells;             // padded containers per buffer
    int32_t *counts[2];
    char *parts[2];
    int cur;
    int peer_valid[2];
    int *overflow;              // device flag: a container overflowed
    char *staging;              // bulk I/O staging (device)
    size_t staging_bytes;
    uint64_t sweeps;

    int64_t cell_index(int x, int y, int z) const { return ((int64_t)(z + 1) * (d[1] + 2) + (y + 1)) * (d[0] + 2) + (x + 1); }
    size_t cell_bytes() const { return (size_t)6 * cap * real; }
};

namespace b200geo {

namespace {

struct BoxDims {
    int nx, ny, nz, cap;
    int org[3];
    double edge;
};

__device__ __forceinline__ int64_t pcell(const BoxDims& D, int x, int y, int z)
{
    return ((int64_t)(z + 1) * (D.ny + 2) + (y + 1)) * (D.nx + 2) + (x + 1);
}

// addContainedParticles over the Moore hood (boxcell.h:123-138,164-174); rebin = 0 copies the
// container (nanoStep != 0).
template<typename REAL>
__global__ void __launch_bounds__(128)
rebin_kernel(const int32_t *__restrict__ cnt_old, const REAL *__restrict__ part_old, int32_t *__restrict__ cnt_new,
             REAL *__restrict__ part_new, BoxDims D, int z0, int z1, int rebin, int *overflow)
{
    const int lane = threadIdx.x & 31;
    const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t row_cells = (int64_t)D.nx * D.ny;
    if (w >= row_cells * (z1 - z0)) return;
    const int z = z0 + (int)(w / row_cells);
    const int y = (int)((w % row_cells) / D.nx), x = (int)(w % D.nx);
    const int64_t self = pcell(D, x, y, z);
    const int cap = D.cap;
    REAL *mine = part_new + self * 6 * cap;
    if (!rebin) {
        int n = cnt_old[self];
        const REAL *src = part_old + self * 6 * cap;
        for (int i = lane; i < 6 * cap; i += 32) mine[i] = src[i];
        if (lane == 0) cnt_new[self] = n;
        return;
    }
    const double ox = (double)(x + D.org[0]) * D.edge, oy = (double)(y + D.org[1]) * D.edge, oz = (double)(z + D.org[2]) * D.edge;
    const double qx = ox + D.edge, qy = oy + D.edge, qz = oz + D.edge;
    int n = 0;
    for (int k = 0; k < 27; ++k) {
        const int64_t nb = pcell(D, x + k % 3 - 1, y + (k / 3) % 3 - 1, z + k / 9 - 1);
        const int c = cnt_old[nb];
        const REAL *src = part_old + nb * 6 * cap;
        for (int base = 0; base < c; base += 32) {
            const int s = base + lane;
            bool inside = false;
            if (s < c) {
                double px = src[s], py = src[cap + s], pz = src[2 * cap + s];
                inside = ox <= px && oy <= py && oz <= pz && px < qx && py < qy && pz < qz;
            }
            const unsigned mask = __ballot_sync(0xffffffffu, inside);
            if (inside) {
                const int dst = n + __popc(mask & ((1u << lane) - 1));
                if (dst < cap) {
#pragma unroll
                    for (int q = 0; q < 6; ++q) mine[q * cap + dst] = src[q * cap + s];
                }
            }
            n += __popc(mask);
        }
    }
    if (lane == 0) {
        cnt_new[self] = n < cap ? n : cap;
        if (n > cap) atomicOr(overflow, 1);  // FixedArray::operator<<: "capacity exceeded"
    }
}

// updateCargo (boxcell.h:140-152) with LJParticle::update (oracle/models/nbody.h).
template<typename REAL, int G>
__global__ void __launch_bounds__(128)
force_kernel(const int32_t *__restrict__ cnt_old, const REAL *__restrict__ part_old, const int32_t *__restrict__ cnt_new,
             REAL *__restrict__ part_new, BoxDims D, int z0, int runs_per_row, REAL dt, REAL rc2)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int cap = D.cap;
    REAL *spos = reinterpret_cast<REAL *>(smem_raw);            // [(G + 2) * 9][3][cap]
    int *scnt = reinterpret_cast<int *>(spos + (G + 2) * 9 * 3 * cap);  // [(G + 2) * 9]
    int *pre = scnt + (G + 2) * 9;                              // [G + 1]

    const int run = blockIdx.x % runs_per_row;
    const int y = (blockIdx.x / runs_per_row) % D.ny;
    const int z = z0 + blockIdx.x / (runs_per_row * D.ny);
    const int x0 = run * G;

    // stage the old neighbourhood of the whole run: local (lx, ly, lz) = cell (x0 - 1 + lx, y - 1 + ly, z - 1 + lz)
    for (int c = threadIdx.x; c < (G + 2) * 9; c += blockDim.x) {
        int lx = c % (G + 2), ly = (c / (G + 2)) % 3, lz = c / ((G + 2) * 3);
        int cx = x0 - 1 + lx;
        scnt[c] = cx <= D.nx ? cnt_old[pcell(D, cx, y - 1 + ly, z - 1 + lz)] : 0;
    }
    if (threadIdx.x <= G) {
        int acc = 0;
        for (int j = 0; j < (int)threadIdx.x; ++j) acc += (x0 + j < D.nx) ? cnt_new[pcell(D, x0 + j, y, z)] : 0;
        pre[threadIdx.x] = acc;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < (G + 2) * 9 * 3 * cap; i += blockDim.x) {
        int c = i / (3 * cap), r = i % (3 * cap);
        if (r % cap < scnt[c]) {
            int lx = c % (G + 2), ly = (c / (G + 2)) % 3, lz = c / ((G + 2) * 3);
            spos[i] = part_old[pcell(D, x0 - 1 + lx, y - 1 + ly, z - 1 + lz) * 6 * cap + r];
        }
    }
    __syncthreads();

    const int total = pre[G];
    for (int t = threadIdx.x; t < total; t += blockDim.x) {
        int j = 0;
        while (pre[j + 1] <= t) ++j;
        const int slot = t - pre[j];
        REAL *me = part_new + pcell(D, x0 + j, y, z) * 6 * cap + slot;
        const REAL p0 = me[0], p1 = me[cap], p2 = me[2 * cap];
        REAL v0 = me[3 * cap], v1 = me[4 * cap], v2 = me[5 * cap];
        for (int k = 0; k < 27; ++k) {
            const int c = ((k / 9) * 3 + (k / 3) % 3) * (G + 2) + j + k % 3;
            const int n = scnt[c];
            const REAL *s = spos + c * 3 * cap;
            for (int p = 0; p < n; ++p) {
                const REAL d0 = p0 - s[p], d1 = p1 - s[cap + p], d2 = p2 - s[2 * cap + p];
                const REAL r2 = (d0 * d0 + d1 * d1) + d2 * d2;
                if (r2 == (REAL)0 || r2 >= rc2) continue;
                const REAL inv = (REAL)1 / r2;
                const REAL s6 = inv * inv * inv;
                const REAL f = ((REAL)24 * inv) * s6 * ((REAL)2 * s6 - (REAL)1);
                v0 += (d0 * f) * dt;
                v1 += (d1 * f) * dt;
                v2 += (d2 * f) * dt;
            }
        }
        me[0] = p0 + v0 * dt;
        me[cap] = p1 + v1 * dt;
        me[2 * cap] = p2 + v2 * dt;
        me[3 * cap] = v0;
        me[4 * cap] = v1;
        me[5 * cap] = v2;
    }
}

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

// one contiguous block global -> shared through the TMA engine, completion counted on an mbarrier
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

template<typename REAL> struct alignas(2 * sizeof(REAL)) Pos2 { REAL x, y; };

__device__ __forceinline__ float fma_any(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ double fma_any(double a, double b, double c) { return __fma_rn(a, b, c); }

// smallest REAL >= v: for a REAL-valued p, (double)p >= v <=> p >= round_up(v) and (double)p < v <=> p < round_up(v),
// so the reference's double-precision position check can be done on the particles' own type
__device__ __forceinline__ void round_up(double v, float& r) { r = __double2float_ru(v); }
__device__ __forceinline__ void round_up(double v, double& r) { r = v; }

// Re-bin and update in ONE kernel: the staged neighbourhood serves both, and the new containers
// are written once.
template<typename REAL, int G, int NT, int LMAX>
__global__ void __launch_bounds__(NT)
fused_kernel(const int32_t *__restrict__ cnt_old, const REAL *__restrict__ part_old, int32_t *__restrict__ cnt_new,
             REAL *__restrict__ part_new, BoxDims D, int z0, int runs_per_row, REAL dt, REAL rc2, REAL rc2_loose,
             int rebin, int *overflow)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int cap = D.cap;
    // a staged container = the position block of the old container as it lies in HBM, x[cap] y[cap]
    // z[cap] (one bulk copy), `SP` elements apart: the 4 extra elements shift consecutive containers
    // by 4 banks, so lanes working on neighbouring containers do not collide in shared memory
    const int SP = 3 * cap + 4;
    constexpr int NC = (G + 2) * 9;
    REAL *spos = reinterpret_cast<REAL *>(smem_raw);                 // [NC][SP]
    int *scnt = reinterpret_cast<int *>(spos + NC * SP);             // [NC]
    int *pre = scnt + NC;                                            // [G + 1]
    uint64_t *bar = reinterpret_cast<uint64_t *>(pre + G + 1 + ((G + 1) & 1));
    unsigned short *newidx = reinterpret_cast<unsigned short *>(bar + 1);   // [G][cap]: staged index of every new particle
    unsigned short *list = newidx + G * cap + ((G * cap) & 1);              // [LMAX][NT]

    const int run = blockIdx.x % runs_per_row;
    const int y = (blockIdx.x / runs_per_row) % D.ny;
    const int z = z0 + blockIdx.x / (runs_per_row * D.ny);
    const int x0 = run * G;

    // stage the old neighbourhood of the whole run: local (lx, ly, lz) = container (x0 - 1 + lx, y - 1 + ly,
    // z - 1 + lz). One thread per container reads its count and, if it holds particles, issues ONE
    // bulk asynchronous copy (cp.async.bulk, the TMA engine) of its 3 * cap positions; all copies are in
    // flight at once and complete on one mbarrier.
    if (threadIdx.x == 0) {
        mbar_init(bar, NC);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    for (int c = threadIdx.x; c < NC; c += NT) {
        int lx = c % (G + 2), ly = (c / (G + 2)) % 3, lz = c / ((G + 2) * 3);
        int cx = x0 - 1 + lx;
        int n = 0;
        if (cx <= D.nx) {
            const int64_t pc = pcell(D, cx, y - 1 + ly, z - 1 + lz);
            n = cnt_old[pc];
            if (n > 0) {
                const uint32_t bytes = 3 * cap * sizeof(REAL);
                mbar_expect_tx(bar, bytes);
                bulk_load(spos + c * SP, part_old + pc * 6 * cap, bytes, bar);
            }
        }
        if (n <= 0) mbar_arrive(bar);
        scnt[c] = n;
    }
    __syncthreads();
    while (!mbar_try_wait(bar, 0)) {}

    // re-bin (boxcell.h:123-138,164-174): one warp per container of the run scans the 27 staged OLD
    // containers in reference order and keeps, by ordered ballot/popc compaction, the particles inside
    // its box; nanoStep != 0 keeps the container as it is
    {
        const int lane = threadIdx.x & 31;
        for (int j = threadIdx.x >> 5; j < G; j += NT / 32) {
            int n = 0;
            if (x0 + j < D.nx) {
                const double ex = (double)(x0 + j + D.org[0]) * D.edge, ey = (double)(y + D.org[1]) * D.edge,
                             ez = (double)(z + D.org[2]) * D.edge;
                REAL ox, oy, oz, qx, qy, qz;
                round_up(ex, ox);
                round_up(ey, oy);
                round_up(ez, oz);
                round_up(ex + D.edge, qx);
                round_up(ey + D.edge, qy);
                round_up(ez + D.edge, qz);
                // the 27 source containers in reference order, TWO per warp iteration where both hold at most 16
                // particles (lanes 0-15 the first, lanes 16-31 the second: the ordered compaction by ballot / popc keeps
                // the order) — containers hold ~13 particles, so a scan of one container kept 13 of 32 lanes busy
                auto scan = [&](int cell, int sl, bool use) {
                    bool inside = false;
                    if (use) {
                        const REAL *q = spos + cell * SP + sl;
                        const REAL px = q[0], py = q[cap], pz = q[2 * cap];
                        inside = !rebin || (ox <= px && oy <= py && oz <= pz && px < qx && py < qy && pz < qz);
                    }
                    const unsigned mask = __ballot_sync(0xffffffffu, inside);
                    if (inside) {
                        const int dst = n + __popc(mask & ((1u << lane) - 1));
                        if (dst < cap) newidx[j * cap + dst] = (unsigned short)(cell * SP + sl);
                    }
                    n += __popc(mask);
                };
                auto cell_of = [&](int k) { return ((k / 9) * 3 + (k / 3) % 3) * (G + 2) + j + k % 3; };
                if (!rebin) {
                    const int cell = cell_of(13), cn = scnt[cell];
                    for (int b = 0; b < cn; b += 32) scan(cell, b + lane, b + lane < cn);
                } else {
#pragma unroll 1
                    for (int k = 0; k < 27; k += 2) {
                        const int c0 = cell_of(k), n0 = scnt[c0];
                        const int c1 = k + 1 < 27 ? cell_of(k + 1) : c0, n1 = k + 1 < 27 ? scnt[c1] : 0;
                        if (n0 <= 16 && n1 <= 16) {
                            const int half = lane >> 4, sl = lane & 15;
                            scan(half ? c1 : c0, sl, sl < (half ? n1 : n0));
                        } else {
                            for (int b = 0; b < n0; b += 32) scan(c0, b + lane, b + lane < n0);
                            for (int b = 0; b < n1; b += 32) scan(c1, b + lane, b + lane < n1);
                        }
                    }
                }
                if (lane == 0) {
                    cnt_new[pcell(D, x0 + j, y, z)] = n < cap ? n : cap;
                    if (n > cap) atomicOr(overflow, 1);  // FixedArray::operator<<: "capacity exceeded"
                }
            }
            if (lane == 0) pre[j + 1] = n < cap ? n : cap;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        pre[0] = 0;
        for (int j = 0; j < G; ++j) pre[j + 1] += pre[j];
    }
    __syncthreads();

    // updateCargo (boxcell.h:140-152) with LJParticle::update, two passes per particle so that the
    // expensive part runs with full lanes:
    //   pass 1 walks the 27 containers in reference order and appends to a per-thread list in shared
    //          memory the candidates inside a slightly ENLARGED cutoff (cheap fused distance; the margin
    //          of 2^-16 covers its rounding difference to the reference expression many times over: it
    //          only decides what pass 2 looks at). Branch-free: every candidate is stored at the list's
    //          end, the end only advances for the accepted ones.
    //   pass 2 walks the list in the same order with the reference arithmetic, including the exact
    //          `r2 == 0 || r2 >= rc2` test, so the result is bit-identical to the one-pass kernel.
    // Only ~15 % of the candidates interact (4/3 pi rc^3 of the 27 rc^3 searched); in a one-pass loop
    // nearly every warp iteration has SOME lane inside the cutoff and all lanes pay for the force.
    const int total = pre[G];
    unsigned short *mylist = list + threadIdx.x;
    for (int t = threadIdx.x; t < total; t += NT) {
        int j = 0;
        while (pre[j + 1] <= t) ++j;
        const int slot = t - pre[j];
        // the particle's old state: position from the staged copy, velocity from its old container
        const int from = newidx[j * cap + slot], fc = from / SP, fs = from % SP;
        const REAL *old = part_old + pcell(D, x0 - 1 + fc % (G + 2), y - 1 + (fc / (G + 2)) % 3, z - 1 + fc / ((G + 2) * 3)) * 6 * cap + fs;
        const REAL p0 = spos[from], p1 = spos[from + cap], p2 = spos[from + 2 * cap];
        REAL v0 = old[3 * cap], v1 = old[4 * cap], v2 = old[5 * cap];

        auto drain = [&](int n) {
            for (int i = 0; i < n; ++i) {
                const REAL *q = spos + mylist[i * NT];
                const REAL d0 = p0 - q[0], d1 = p1 - q[cap], d2 = p2 - q[2 * cap];
                const REAL r2 = (d0 * d0 + d1 * d1) + d2 * d2;
                if (r2 == (REAL)0 || r2 >= rc2) continue;
                const REAL inv = (REAL)1 / r2;
                const REAL s6 = inv * inv * inv;
                const REAL f = ((REAL)24 * inv) * s6 * ((REAL)2 * s6 - (REAL)1);
                v0 += (d0 * f) * dt;
                v1 += (d1 * f) * dt;
                v2 += (d2 * f) * dt;
            }
        };

        unsigned short *lp = mylist;
        for (int kk = 0; kk < 9; ++kk) {
            for (int kx = 0; kx < 3; ++kx) {
                const int cell = kk * (G + 2) + j + kx;
                const int cn = scnt[cell];
                if ((int)(lp - mylist) + cn * NT > LMAX * NT) {  // the list could overflow: evaluate what is there
                    drain((int)(lp - mylist) / NT);
                    lp = mylist;
                }
                const REAL *q = spos + cell * SP;
                const int idx0 = cell * SP;
#pragma unroll 4
                for (int p = 0; p < cn; ++p) {
                    const REAL d0 = p0 - q[p], d1 = p1 - q[cap + p], d2 = p2 - q[2 * cap + p];
                    const REAL r2 = fma_any(d2, d2, fma_any(d1, d1, d0 * d0));
                    *lp = (unsigned short)(idx0 + p);
                    lp += r2 < rc2_loose ? NT : 0;
                }
            }
        }
        drain((int)(lp - mylist) / NT);

        REAL *me = part_new + pcell(D, x0 + j, y, z) * 6 * cap + slot;
        me[0] = p0 + v0 * dt;
        me[cap] = p1 + v1 * dt;
        me[2 * cap] = p2 + v2 * dt;
        me[3 * cap] = v0;
        me[4 * cap] = v1;
        me[5 * cap] = v2;
    }
}

// 128-bit shared-memory load of V = 16 / sizeof(REAL) consecutive coordinates
__device__ __forceinline__ void lds_vec(const float *p, float (&v)[4])
{
    const float4 t = *reinterpret_cast<const float4 *>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void lds_vec(const double *p, double (&v)[2])
{
    const double2 t = *reinterpret_cast<const double2 *>(p);
    v[0] = t.x; v[1] = t.y;
}

// The fused re-bin / update kernel, second generation (capacity <= 32). Same staging and re-bin as fused_kernel;
// the particle update differs:
//   pass 1 (filter): per neighbour container ONE 32-bit mask of the slots inside the enlarged cutoff, built from
//          128-bit shared-memory loads (4 floats / 2 doubles of x, of y, of z per load: 0.75 loads per candidate
//          instead of 3) and a fused distance; the 27 masks go to shared memory ([27][NT] words: 27 stores per
//          particle where the candidate lists took one store per CANDIDATE, and 27.6 KB per CTA instead of 45 KB);
//   pass 2 (exact): walks the set bits of the 27 masks in order — container by container, slot by slot: the
//          reference's order — as ONE flat loop per lane (a lane refills its mask word when it runs dry, lanes do
//          not wait for each other per container), with the reference arithmetic incl. the exact cutoff test.
template<typename REAL, int G, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB)
fused2_kernel(const int32_t *__restrict__ cnt_old, const REAL *__restrict__ part_old, int32_t *__restrict__ cnt_new,
              REAL *__restrict__ part_new, BoxDims D, int z0, int runs_per_row, REAL dt, REAL rc2, REAL rc2_loose,
              int rebin, int *overflow)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int V = 16 / (int)sizeof(REAL);
    const int cap = D.cap;
    const int SP = 3 * cap + 4;      // see fused_kernel
    constexpr int NC = (G + 2) * 9;
    REAL *spos = reinterpret_cast<REAL *>(smem_raw);                 // [NC][SP]
    uint32_t *smask = reinterpret_cast<uint32_t *>(spos + NC * SP);  // [27][NT]
    int *scnt = reinterpret_cast<int *>(smask + 27 * NT);            // [NC]
    int *pre = scnt + NC;                                            // [G + 1]
    uint64_t *bar = reinterpret_cast<uint64_t *>(pre + G + 1 + ((G + 1) & 1));
    unsigned short *newidx = reinterpret_cast<unsigned short *>(bar + 1);   // [G][cap]: staged index of every new particle

    const int run = blockIdx.x % runs_per_row;
    const int y = (blockIdx.x / runs_per_row) % D.ny;
    const int z = z0 + blockIdx.x / (runs_per_row * D.ny);
    const int x0 = run * G;

    if (threadIdx.x == 0) {
        mbar_init(bar, NC);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    for (int c = threadIdx.x; c < NC; c += NT) {
        int lx = c % (G + 2), ly = (c / (G + 2)) % 3, lz = c / ((G + 2) * 3);
        int cx = x0 - 1 + lx;
        int n = 0;
        if (cx <= D.nx) {
            const int64_t pc = pcell(D, cx, y - 1 + ly, z - 1 + lz);
            n = cnt_old[pc];
            if (n > 0) {
                const uint32_t bytes = 3 * cap * sizeof(REAL);
                mbar_expect_tx(bar, bytes);
                bulk_load(spos + c * SP, part_old + pc * 6 * cap, bytes, bar);
            }
        }
        if (n <= 0) mbar_arrive(bar);
        scnt[c] = n;
    }
    __syncthreads();
    while (!mbar_try_wait(bar, 0)) {}

    // re-bin: as in fused_kernel (boxcell.h:123-138,164-174)
    {
        const int lane = threadIdx.x & 31;
        for (int j = threadIdx.x >> 5; j < G; j += NT / 32) {
            int n = 0;
            if (x0 + j < D.nx) {
                const double ex = (double)(x0 + j + D.org[0]) * D.edge, ey = (double)(y + D.org[1]) * D.edge,
                             ez = (double)(z + D.org[2]) * D.edge;
                REAL ox, oy, oz, qx, qy, qz;
                round_up(ex, ox);
                round_up(ey, oy);
                round_up(ez, oz);
                round_up(ex + D.edge, qx);
                round_up(ey + D.edge, qy);
                round_up(ez + D.edge, qz);
#pragma unroll
                for (int k = 0; k < 27; ++k) {
                    if (!rebin && k != 13) continue;
                    const int cell = ((k / 9) * 3 + (k / 3) % 3) * (G + 2) + j + k % 3;
                    const int cn = scnt[cell];
                    for (int b = 0; b < cn; b += 32) {
                        const int sl = b + lane;
                        bool inside = false;
                        if (sl < cn) {
                            const REAL *q = spos + cell * SP + sl;
                            const REAL px = q[0], py = q[cap], pz = q[2 * cap];
                            inside = !rebin || (ox <= px && oy <= py && oz <= pz && px < qx && py < qy && pz < qz);
                        }
                        const unsigned mask = __ballot_sync(0xffffffffu, inside);
                        if (inside) {
                            const int dst = n + __popc(mask & ((1u << lane) - 1));
                            if (dst < cap) newidx[j * cap + dst] = (unsigned short)(cell * SP + sl);
                        }
                        n += __popc(mask);
                    }
                }
                if (lane == 0) {
                    cnt_new[pcell(D, x0 + j, y, z)] = n < cap ? n : cap;
                    if (n > cap) atomicOr(overflow, 1);  // FixedArray::operator<<: "capacity exceeded"
                }
            }
            if (lane == 0) pre[j + 1] = n < cap ? n : cap;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        pre[0] = 0;
        for (int j = 0; j < G; ++j) pre[j + 1] += pre[j];
    }
    __syncthreads();

    const int total = pre[G];
    uint32_t *mymask = smask + threadIdx.x;
    for (int t = threadIdx.x; t < total; t += NT) {
        int j = 0;
        while (pre[j + 1] <= t) ++j;
        const int slot = t - pre[j];
        const int from = newidx[j * cap + slot], fc = from / SP, fs = from % SP;
        const REAL *old = part_old + pcell(D, x0 - 1 + fc % (G + 2), y - 1 + (fc / (G + 2)) % 3, z - 1 + fc / ((G + 2) * 3)) * 6 * cap + fs;
        const REAL p0 = spos[from], p1 = spos[from + cap], p2 = spos[from + 2 * cap];
        REAL v0 = old[3 * cap], v1 = old[4 * cap], v2 = old[5 * cap];

        // pass 1: one mask per neighbour container (the row loop is NOT unrolled: 27 copies of the candidate loop
        // thrash the instruction cache — 24 % of the stall samples were no_inst, profiles/r3d)
#pragma unroll 1
        for (int kk = 0; kk < 9; ++kk) {
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int cell = kk * (G + 2) + j + kx;
                const int cn = scnt[cell];
                const REAL *q = spos + cell * SP;
                uint32_t m = 0;
#pragma unroll 2
                for (int p = 0; p < cn; p += V) {
                    REAL xs[V], ys[V], zs[V];
                    lds_vec(q + p, xs);
                    lds_vec(q + cap + p, ys);
                    lds_vec(q + 2 * cap + p, zs);
                    uint32_t nib = 0;
#pragma unroll
                    for (int u = 0; u < V; ++u) {
                        const REAL d0 = p0 - xs[u], d1 = p1 - ys[u], d2 = p2 - zs[u];
                        const REAL r2 = fma_any(d2, d2, fma_any(d1, d1, d0 * d0));
                        if (r2 < rc2_loose) nib |= 1u << u;
                    }
                    m |= nib << p;
                }
                // slots at and beyond cn of the last group hold stale shared memory
                m &= cn >= 32 ? 0xffffffffu : (1u << cn) - 1u;
                mymask[(kk * 3 + kx) * NT] = m;
            }
        }

        // pass 2: the accepted candidates in reference order, reference arithmetic
        {
            int k = -1;
            uint32_t m = 0;
            const REAL *q = spos;
            for (;;) {
                while (m == 0 && k < 26) {
                    ++k;
                    m = mymask[k * NT];
                    q = spos + ((k / 3) * (G + 2) + j + k % 3) * SP;
                }
                if (m == 0) break;
                const int p = __ffs(m) - 1;
                m &= m - 1;
                const REAL d0 = p0 - q[p], d1 = p1 - q[cap + p], d2 = p2 - q[2 * cap + p];
                const REAL r2 = (d0 * d0 + d1 * d1) + d2 * d2;
                if (r2 == (REAL)0 || r2 >= rc2) continue;
                const REAL inv = (REAL)1 / r2;
                const REAL s6 = inv * inv * inv;
                const REAL f = ((REAL)24 * inv) * s6 * ((REAL)2 * s6 - (REAL)1);
                v0 += (d0 * f) * dt;
                v1 += (d1 * f) * dt;
                v2 += (d2 * f) * dt;
            }
        }

        REAL *me = part_new + pcell(D, x0 + j, y, z) * 6 * cap + slot;
        me[0] = p0 + v0 * dt;
        me[cap] = p1 + v1 * dt;
        me[2 * cap] = p2 + v2 * dt;
        me[3 * cap] = v0;
        me[4 * cap] = v1;
        me[5 * cap] = v2;
    }
}

// dense AoS [cells][cap][6] (host interchange format) <-> SoA containers of a box
template<typename REAL, bool LOAD>
__global__ void transpose_kernel(int32_t *cnt, REAL *part, BoxDims D, int ox, int oy, int oz, int dx, int dy, int dz,
                                 int32_t *dense_cnt, REAL *dense_part)
{
    const int64_t cells = (int64_t)dx * dy * dz;
    const int cap = D.cap;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < cells * cap * 6; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t c = i / (cap * 6);
        int r = (int)(i % (cap * 6)), s = r / 6, q = r % 6;
        int x = (int)(c % dx), y = (int)((c / dx) % dy), z = (int)(c / ((int64_t)dx * dy));
        int64_t pc = pcell(D, ox + x, oy + y, oz + z);
        if (LOAD) {
            part[pc * 6 * cap + q * cap + s] = dense_part[i];
            if (r == 0) cnt[pc] = dense_cnt[c];
        } else {
            dense_part[i] = s < cnt[pc] ? part[pc * 6 * cap + q * cap + s] : (REAL)0;
            if (r == 0) dense_cnt[c] = cnt[pc];
        }
    }
}

BoxDims box_dims(const b200geo_boxgrid *g)
{
    BoxDims D;
    D.nx = g->d[0];
    D.ny = g->d[1];
    D.nz = g->d[2];
    D.cap = g->cap;
    for (int i = 0; i < 3; ++i) D.org[i] = g->desc.cell_origin[i];
    D.edge = g->desc.cell_edge;
    return D;
}

constexpr int RUN = 8;  // containers per CTA of the force kernel

template<typename REAL>
int sweep(b200geo_boxgrid *g, const b200geo_nbody_params *p, int rebin, cudaStream_t s)
{
    BoxDims D = box_dims(g);
    const int32_t *co = g->counts[g->cur];
    int32_t *cn = g->counts[g->cur ^ 1];
    const REAL *po = (const REAL *)g->parts[g->cur];
    REAL *pn = (REAL *)g->parts[g->cur ^ 1];
    int64_t cells = (int64_t)D.nx * D.ny * D.nz;
    REAL rc = (REAL)p->cutoff;
    if (g_tuning.nbody_kernel == 1 || g->cap % 4 != 0) {
        // two kernels: re-bin, then the one-pass force kernel (any capacity)
        rebin_kernel<REAL><<<(unsigned)((cells + 3) / 4), 128, 0, s>>>(co, po, cn, pn, D, 0, D.nz, rebin, g->overflow);
        count_launch();
        int runs = (D.nx + RUN - 1) / RUN;
        size_t smem = (size_t)(RUN + 2) * 9 * 3 * g->cap * sizeof(REAL) + ((RUN + 2) * 9 + RUN + 1) * sizeof(int);
        static bool attr[64] = {false};   // per device
        if (!attr[g->device & 63]) {
            B200GEO_CUDA(cudaFuncSetAttribute(force_kernel<REAL, RUN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
            attr[g->device & 63] = true;
        }
        force_kernel<REAL, RUN><<<(unsigned)((int64_t)runs * D.ny * D.nz), 128, smem, s>>>(co, po, cn, pn, D, 0, runs, (REAL)p->dt, rc * rc);
    } else if (g_tuning.nbody_kernel == 3 || g->cap > 32) {
        // first-generation fused kernel (per-thread candidate lists): any capacity up to 64. Runs of G containers hold
        // ~13 G particles on 16 G threads; fewer threads ("nbody.threads") keep more lanes busy but send the runs that hold
        // more particles than threads through the whole loop twice, and lose
        constexpr int G = sizeof(REAL) == 4 ? 16 : 8, LMAX = 88, NC = (G + 2) * 9;
        REAL loose = rc * rc * (REAL)(1.0 + 1.0 / 65536.0);
        int runs = (D.nx + G - 1) / G;
#define NBODY_LAUNCH1(NT)                                                                                                   \
        do {                                                                                                                \
            size_t smem = (size_t)NC * (3 * g->cap + 4) * sizeof(REAL) + (NC + G + 2) * sizeof(int) + 8 +                   \
                          (size_t)(G * g->cap + 2) * sizeof(unsigned short) + (size_t)LMAX * NT * sizeof(unsigned short);   \
            if (smem > 227 * 1024) return fail(B200GEO_ERR_LOGIC, "container capacity too large for the fused kernel's shared memory"); \
            static bool attr3[64] = {false};                                                                                \
            if (!attr3[g->device & 63]) {                                                                                   \
                B200GEO_CUDA(cudaFuncSetAttribute(fused_kernel<REAL, G, NT, LMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
                attr3[g->device & 63] = true;                                                                               \
            }                                                                                                               \
            fused_kernel<REAL, G, NT, LMAX><<<(unsigned)((int64_t)runs * D.ny * D.nz), NT, smem, s>>>(                       \
                co, po, cn, pn, D, 0, runs, (REAL)p->dt, rc * rc, loose, rebin, g->overflow);                               \
        } while (0)
        // (whole warps only: 14 per container for runs of 16, i.e. float)
        constexpr bool LEAN = (14 * G) % 32 == 0;
        const int threads = g_tuning.nbody_threads > 0 ? g_tuning.nbody_threads : 16 * G;   // measured: 256 / 224 / 192 threads = 14.1 / 16.2 / 17.0 ms (profiles/r3j)
        if constexpr (LEAN) {
            if (threads == 14 * G) NBODY_LAUNCH1(14 * G);
            else if (threads == 12 * G) NBODY_LAUNCH1(12 * G);
            else NBODY_LAUNCH1(16 * G);
        } else {
            NBODY_LAUNCH1(16 * G);
        }
#undef NBODY_LAUNCH1
    } else {
        // second-generation fused kernel (per-container masks); run length / CTA size: "nbody.run" = 8 or 16
        REAL loose = rc * rc * (REAL)(1.0 + 1.0 / 65536.0);
#define NBODY_LAUNCH2(G, NT, MINB)                                                                                          \
        do {                                                                                                                \
            constexpr int NC = (G + 2) * 9;                                                                                 \
            int runs = (D.nx + G - 1) / G;                                                                                  \
            size_t smem = (size_t)NC * (3 * g->cap + 4) * sizeof(REAL) + (size_t)27 * NT * 4 + (NC + G + 2) * sizeof(int) + 8 +  \
                          (size_t)(G * g->cap + 2) * sizeof(unsigned short);                                                \
            static bool attr4[64] = {false};                                                                                \
            if (!attr4[g->device & 63]) {                                                                                   \
                B200GEO_CUDA(cudaFuncSetAttribute(fused2_kernel<REAL, G, NT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
                attr4[g->device & 63] = true;                                                                               \
            }                                                                                                               \
            fused2_kernel<REAL, G, NT, MINB><<<(unsigned)((int64_t)runs * D.ny * D.nz), NT, smem, s>>>(                      \
                co, po, cn, pn, D, 0, runs, (REAL)p->dt, rc * rc, loose, rebin, g->overflow);                               \
        } while (0)
        // 0 = automatic: runs of 16 for float, of 8 for double (its staged neighbourhood is twice as large)
        const int run = g_tuning.nbody_run > 0 ? g_tuning.nbody_run : (sizeof(REAL) == 4 ? 16 : 8);
        if (run == 8) NBODY_LAUNCH2(8, 128, 4);
        else if (run == 12) NBODY_LAUNCH2(12, 192, 3);
        else NBODY_LAUNCH2(16, 256, 2);
#undef NBODY_LAUNCH2
    }
    count_launch();
    return check_cuda(cudaGetLastError(), "n-body sweep");
}

int ensure_staging(b200geo_boxgrid *g, size_t bytes, cudaStream_t s)
{
    if (g->staging_bytes >= bytes) return 0;
    if (g->staging) {
        cudaStreamSynchronize(s);
        cudaFree(g->staging);
        g->staging = 0;
        g->staging_bytes = 0;
    }
    cudaError_t e = cudaMalloc((void **)&g->staging, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(B200GEO_ERR_NOMEM, std::string("cudaMalloc failed: ") + cudaGetErrorString(e));
    }
    g->staging_bytes = bytes;
    return 0;
}

bool valid_cell_box(const b200geo_boxgrid *g, const int32_t o[3], const int32_t d[3])
{
    for (int i = 0; i < 3; ++i)
        if (d[i] < 0 || o[i] < -1 || o[i] + d[i] > g->d[i] + 1) return false;
    return true;
}

template<typename REAL>
int box_io(b200geo_boxgrid *g, const int32_t o[3], const int32_t d[3], void *counts, void *parts, int location,
           bool load, int both, cudaStream_t s)
{
    int64_t cells = (int64_t)d[0] * d[1] * d[2];
    if (cells == 0) return B200GEO_OK;
    size_t cb = (size_t)cells * sizeof(int32_t), pb = (size_t)cells * g->cap * 6 * sizeof(REAL);
    int32_t *dc = (int32_t *)counts;
    REAL *dp = (REAL *)parts;
    if (location == B200GEO_HOST) {
        size_t off = (cb + 255) / 256 * 256;
        int rc = ensure_staging(g, off + pb, s);
        if (rc) return rc;
        dc = (int32_t *)g->staging;
        dp = (REAL *)(g->staging + off);
        if (load) {
            B200GEO_CUDA(cudaMemcpyAsync(dc, counts, cb, cudaMemcpyHostToDevice, s));
            B200GEO_CUDA(cudaMemcpyAsync(dp, parts, pb, cudaMemcpyHostToDevice, s));
        }
    }
    BoxDims D = box_dims(g);
    int blocks = (int)((cells * g->cap * 6 + 255) / 256 < 148 * 16 ? (cells * g->cap * 6 + 255) / 256 : 148 * 16);
    if (load) {
        for (int b = 0; b < (both ? 2 : 1); ++b) {
            transpose_kernel<REAL, true><<<blocks, 256, 0, s>>>(g->counts[g->cur ^ b], (REAL *)g->parts[g->cur ^ b], D,
                                                               o[0], o[1], o[2], d[0], d[1], d[2], dc, dp);
            count_launch();
        }
    } else {
        transpose_kernel<REAL, false><<<blocks, 256, 0, s>>>(g->counts[g->cur], (REAL *)g->parts[g->cur], D,
                                                            o[0], o[1], o[2], d[0], d[1], d[2], dc, dp);
        count_launch();
        if (location == B200GEO_HOST) {
            B200GEO_CUDA(cudaMemcpyAsync(counts, dc, cb, cudaMemcpyDeviceToHost, s));
            B200GEO_CUDA(cudaMemcpyAsync(parts, dp, pb, cudaMemcpyDeviceToHost, s));
        }
    }
    B200GEO_CUDA(cudaGetLastError());
    if (location == B200GEO_HOST) B200GEO_CUDA(cudaStreamSynchronize(s));  // the staging buffer is reused
    return B200GEO_OK;
}

}

}

using namespace b200geo;

extern "C" {

int b200geo_boxgrid_create(const b200geo_boxgrid_desc *desc, int device, b200geo_boxgrid **out)
{
    if (!desc || !out) return fail(B200GEO_ERR_INVALID, "null argument");
    for (int i = 0; i < 3; ++i) {
        if (desc->dim[i] < 1) return fail(B200GEO_ERR_INVALID, "grid dimension must be >= 1");
        for (int s = 0; s < 2; ++s) {
            int mode = desc->ghost_mode[i][s];
            if (mode == B200GEO_GHOST_WRAP)
                return fail(B200GEO_ERR_LOGIC, "BoxCell grids are Cube grids: the reference's BoxCell does not shift positions across a Torus seam");
            if (mode == B200GEO_GHOST_PEER && i != 2)
                return fail(B200GEO_ERR_LOGIC, "PEER ghost layers are supported on the last axis only (slab partition)");
            if (mode != B200GEO_GHOST_EDGE && mode != B200GEO_GHOST_PEER) return fail(B200GEO_ERR_INVALID, "bad ghost mode");
        }
    }
    if (desc->capacity < 1 || desc->capacity > 64) return fail(B200GEO_ERR_INVALID, "capacity must be 1..64");
    if (desc->real_bytes != 4 && desc->real_bytes != 8) return fail(B200GEO_ERR_INVALID, "real_bytes must be 4 or 8");
    if (!(desc->cell_edge > 0)) return fail(B200GEO_ERR_INVALID, "cell_edge must be positive");
    B200GEO_CUDA(cudaSetDevice(device));
    b200geo_boxgrid *g = new (std::nothrow) b200geo_boxgrid();
    if (!g) return fail(B200GEO_ERR_NOMEM, "out of host memory");
    memset(g, 0, sizeof(*g));
    g->desc = *desc;
    g->device = device;
    for (int i = 0; i < 3; ++i) g->d[i] = desc->dim[i];
    g->cap = desc->capacity;
    g->real = desc->real_bytes;
    g->pcells = (int64_t)(g->d[0] + 2) * (g->d[1] + 2) * (g->d[2] + 2);
    cudaError_t e = cudaSuccess;
    for (int b = 0; b < 2 && e == cudaSuccess; ++b) {
        e = cudaMalloc((void **)&g->counts[b], (size_t)g->pcells * sizeof(int32_t));
        if (e == cudaSuccess) e = cudaMalloc((void **)&g->parts[b], (size_t)g->pcells * g->cell_bytes());
    }
    if (e == cudaSuccess) e = cudaMalloc((void **)&g->overflow, sizeof(int));
    if (e != cudaSuccess) {
        b200geo_boxgrid_destroy(g);
        cudaGetLastError();
        return fail(B200GEO_ERR_NOMEM, std::string("cudaMalloc failed: ") + cudaGetErrorString(e));
    }
    // every container starts empty: the EDGE ring is the reference's default-constructed edge cell
    for (int b = 0; b < 2; ++b) {
        cudaMemset(g->counts[b], 0, (size_t)g->pcells * sizeof(int32_t));
        cudaMemset(g->parts[b], 0, (size_t)g->pcells * g->cell_bytes());
    }
    cudaMemset(g->overflow, 0, sizeof(int));
    *out = g;
    return B200GEO_OK;
}

int b200geo_boxgrid_destroy(b200geo_boxgrid *g)
{
    if (!g) return B200GEO_OK;
    cudaSetDevice(g->device);
    for (int b = 0; b < 2; ++b) {
        if (g->counts[b]) cudaFree(g->counts[b]);
        if (g->parts[b]) cudaFree(g->parts[b]);
    }
    if (g->overflow) cudaFree(g->overflow);
    if (g->staging) cudaFree(g->staging);
    delete g;
    return B200GEO_OK;
}

int b200geo_boxgrid_load(b200geo_boxgrid *g, const int32_t origin[3], const int32_t dim[3], const int32_t *counts,
                         const void *particles, int location, int both, void *stream)
{
    if (!g || !origin || !dim || !counts || !particles) return fail(B200GEO_ERR_INVALID, "null argument");
    if (!valid_cell_box(g, origin, dim)) return fail(B200GEO_ERR_INVALID, "box outside the grid");
    B200GEO_CUDA(cudaSetDevice(g->device));
    if (g->real == 4)
        return box_io<float>(g, origin, dim, const_cast<int32_t *>(counts), const_cast<void *>(particles), location, true, both, (cudaStream_t)stream);
    return box_io<double>(g, origin, dim, const_cast<int32_t *>(counts), const_cast<void *>(particles), location, true, both, (cudaStream_t)stream);
}

int b200geo_boxgrid_save(const b200geo_boxgrid *g, const int32_t origin[3], const int32_t dim[3], int32_t *counts,
                         void *particles, int location, void *stream)
{
    if (!g || !origin || !dim || !counts || !particles) return fail(B200GEO_ERR_INVALID, "null argument");
    if (!valid_cell_box(g, origin, dim)) return fail(B200GEO_ERR_INVALID, "box outside the grid");
    B200GEO_CUDA(cudaSetDevice(g->device));
    b200geo_boxgrid *m = const_cast<b200geo_boxgrid *>(g);
    if (g->real == 4) return box_io<float>(m, origin, dim, counts, particles, location, false, 0, (cudaStream_t)stream);
    return box_io<double>(m, origin, dim, counts, particles, location, false, 0, (cudaStream_t)stream);
}

int b200geo_boxgrid_step(b200geo_boxgrid *g, const b200geo_nbody_params *params, uint32_t first_nano_step,
                         uint32_t n_steps, void *stream)
{
    if (!g || !params) return fail(B200GEO_ERR_INVALID, "null argument");
    if (params->nano_steps < 1) return fail(B200GEO_ERR_INVALID, "nano_steps must be >= 1");
    if (!(params->cutoff <= g->desc.cell_edge))
        return fail(B200GEO_ERR_INVALID, "cutoff larger than the container edge: interactions would be missed");
    B200GEO_CUDA(cudaSetDevice(g->device));
    cudaStream_t s = (cudaStream_t)stream;
    for (uint32_t t = 0; t < n_steps; ++t) {
        for (int side = 0; side < 2; ++side) {
            if (g->desc.ghost_mode[2][side] != B200GEO_GHOST_PEER) continue;
            if (g->peer_valid[side] < 1)
                return fail(B200GEO_ERR_LOGIC, "ghost zone exhausted: exchange halos before stepping");
        }
        int rebin = ((first_nano_step + t) % (uint32_t)params->nano_steps) == 0;
        int rc = g->real == 4 ? sweep<float>(g, params, rebin, s) : sweep<double>(g, params, rebin, s);
        if (rc) return rc;
        g->cur ^= 1;
        for (int side = 0; side < 2; ++side)
            if (g->desc.ghost_mode[2][side] == B200GEO_GHOST_PEER) --g->peer_valid[side];
        ++g->sweeps;
    }
    return B200GEO_OK;
}

int b200geo_boxgrid_check(b200geo_boxgrid *g, void *stream)
{
    if (!g) return fail(B200GEO_ERR_INVALID, "null grid");
    B200GEO_CUDA(cudaSetDevice(g->device));
    int flag = 0;
    B200GEO_CUDA(cudaMemcpyAsync(&flag, g->overflow, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    B200GEO_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    if (flag) {
        cudaMemsetAsync(g->overflow, 0, sizeof(int), (cudaStream_t)stream);
        return fail(B200GEO_ERR_OUT_OF_RANGE, "capacity exceeded");
    }
    return B200GEO_OK;
}

int b200geo_boxgrid_halo_block(const b200geo_boxgrid *g, int array, int side, int kind, int which, void **ptr,
                               uint64_t *bytes)
{
    if (!g || !ptr || !bytes) return fail(B200GEO_ERR_INVALID, "null argument");
    if ((array != 0 && array != 1) || (side != 0 && side != 1) || (kind != 0 && kind != 1) || (which != 0 && which != 1))
        return fail(B200GEO_ERR_INVALID, "bad array/side/kind");
    const int nz = g->d[2];
    int64_t plane = (int64_t)(g->d[0] + 2) * (g->d[1] + 2);
    int64_t zp = kind == 0 ? (side == 0 ? 1 : nz) : (side == 0 ? 0 : nz + 1);  // padded plane index
    int buf = g->cur ^ which;
    if (array == 0) {
        *ptr = g->counts[buf] + zp * plane;
        *bytes = (uint64_t)plane * sizeof(int32_t);
    } else {
        *ptr = g->parts[buf] + zp * plane * g->cell_bytes();
        *bytes = (uint64_t)plane * g->cell_bytes();
    }
    return B200GEO_OK;
}

int b200geo_boxgrid_halo_mark_valid(b200geo_boxgrid *g, int side, int width)
{
    if (!g || (side != 0 && side != 1) || width < 0 || width > 1) return fail(B200GEO_ERR_INVALID, "bad side/width");
    g->peer_valid[side] = width;
    return B200GEO_OK;
}

/* ---- slab groups of container grids: one host thread, several GPUs (twin of csrc/group.cu) ------------- */

}

#define B200GEO_BOXGROUP_MAX 16

struct b200geo_boxgroup {
    int n;
    b200geo_boxgrid *g[B200GEO_BOXGROUP_MAX];
    cudaStream_t stream[B200GEO_BOXGROUP_MAX];
    cudaEvent_t done[B200GEO_BOXGROUP_MAX];
    uint64_t exchanges, bytes_moved;
};

extern "C" {

int b200geo_boxgroup_create(b200geo_boxgrid *const *grids, int n, b200geo_boxgroup **out)
{
    if (!grids || !out || n < 1 || n > B200GEO_BOXGROUP_MAX) return fail(B200GEO_ERR_INVALID, "bad slab group");
    for (int i = 0; i < n; ++i) {
        const b200geo_boxgrid *g = grids[i];
        if (!g) return fail(B200GEO_ERR_INVALID, "null grid in slab group");
        if (g->d[0] != grids[0]->d[0] || g->d[1] != grids[0]->d[1] || g->cap != grids[0]->cap || g->real != grids[0]->real)
            return fail(B200GEO_ERR_INVALID, "slabs of one group must have the same cross-section, capacity and particle type");
        bool low = n > 1 && i > 0, high = n > 1 && i < n - 1;
        if ((g->desc.ghost_mode[2][0] == B200GEO_GHOST_PEER) != low || (g->desc.ghost_mode[2][1] == B200GEO_GHOST_PEER) != high)
            return fail(B200GEO_ERR_INVALID, "slab faces towards a neighbour must be PEER ghost layers, outer faces must not");
    }
    b200geo_boxgroup *grp = new (std::nothrow) b200geo_boxgroup();
    if (!grp) return fail(B200GEO_ERR_NOMEM, "out of host memory");
    memset(grp, 0, sizeof(*grp));
    grp->n = n;
    for (int i = 0; i < n; ++i) grp->g[i] = grids[i];
    for (int i = 0; i < n; ++i) {
        int dev = grids[i]->device;
        cudaError_t e = cudaSetDevice(dev);
        for (int j = i - 1; j <= i + 1 && e == cudaSuccess; j += 2) {
            if (j < 0 || j >= n || grids[j]->device == dev) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, dev, grids[j]->device) == cudaSuccess && can) {
                cudaError_t p = cudaDeviceEnablePeerAccess(grids[j]->device, 0);
                if (p != cudaSuccess && p != cudaErrorPeerAccessAlreadyEnabled) e = p;
                cudaGetLastError();
            }
        }
        if (e == cudaSuccess) e = cudaStreamCreate(&grp->stream[i]);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&grp->done[i], cudaEventDisableTiming);
        if (e != cudaSuccess) {
            b200geo_boxgroup_destroy(grp);
            return check_cuda(e, "slab group setup");
        }
    }
    *out = grp;
    return B200GEO_OK;
}

int b200geo_boxgroup_destroy(b200geo_boxgroup *grp)
{
    if (!grp) return B200GEO_OK;
    for (int i = 0; i < grp->n; ++i) {
        cudaSetDevice(grp->g[i]->device);
        if (grp->stream[i]) { cudaStreamSynchronize(grp->stream[i]); cudaStreamDestroy(grp->stream[i]); }
        if (grp->done[i]) cudaEventDestroy(grp->done[i]);
    }
    delete grp;
    return B200GEO_OK;
}

/* n_steps x { pull the neighbours' boundary container planes (counts and particles, one contiguous block
 * each) into the ghost planes over NVLink; re-bin / update every slab; swap }. BoxCell pulls its particles
 * from the neighbourhood, so the ghost plane IS the migration message (storage/boxcell.h:123-138). */
int b200geo_boxgroup_step(b200geo_boxgroup *grp, const b200geo_nbody_params *params, uint32_t first_nano_step, uint32_t n_steps)
{
    if (!grp || !params) return fail(B200GEO_ERR_INVALID, "null argument");
    for (uint32_t t = 0; t < n_steps; ++t) {
        if (grp->n > 1) {
            for (int i = 0; i < grp->n; ++i) {
                B200GEO_CUDA(cudaSetDevice(grp->g[i]->device));
                B200GEO_CUDA(cudaEventRecord(grp->done[i], grp->stream[i]));
            }
            for (int i = 0; i < grp->n; ++i) {
                b200geo_boxgrid *g = grp->g[i];
                B200GEO_CUDA(cudaSetDevice(g->device));
                for (int side = 0; side < 2; ++side) {
                    int j = side == 0 ? i - 1 : i + 1;
                    if (j < 0 || j >= grp->n) continue;
                    b200geo_boxgrid *peer = grp->g[j];
                    B200GEO_CUDA(cudaStreamWaitEvent(grp->stream[i], grp->done[j], 0));
                    for (int array = 0; array < 2; ++array) {
                        void *src = 0, *dst = 0;
                        uint64_t bytes = 0, dst_bytes = 0;
                        // our low-side ghost plane = the low neighbour's HIGH boundary plane, and vice versa
                        int rc = b200geo_boxgrid_halo_block(peer, array, 1 - side, 0, 0, &src, &bytes);
                        if (rc) return rc;
                        rc = b200geo_boxgrid_halo_block(g, array, side, 1, 0, &dst, &dst_bytes);
                        if (rc) return rc;
                        if (bytes != dst_bytes) return fail(B200GEO_ERR_INVALID, "slabs of one group must have the same cross-section");
                        B200GEO_CUDA(cudaMemcpyPeerAsync(dst, g->device, src, peer->device, bytes, grp->stream[i]));
                        grp->bytes_moved += bytes;
                    }
                    g->peer_valid[side] = 1;
                }
            }
            ++grp->exchanges;
        }
        for (int i = 0; i < grp->n; ++i) {
            int rc = b200geo_boxgrid_step(grp->g[i], params, first_nano_step + t, 1, grp->stream[i]);
            if (rc) return rc;
        }
    }
    return B200GEO_OK;
}

int b200geo_boxgroup_sync(b200geo_boxgroup *grp)
{
    if (!grp) return fail(B200GEO_ERR_INVALID, "null group");
    for (int i = 0; i < grp->n; ++i) {
        B200GEO_CUDA(cudaSetDevice(grp->g[i]->device));
        B200GEO_CUDA(cudaStreamSynchronize(grp->stream[i]));
    }
    return B200GEO_OK;
}

int b200geo_boxgroup_stats(const b200geo_boxgroup *grp, uint64_t out[2])
{
    if (!grp || !out) return fail(B200GEO_ERR_INVALID, "null argument");
    out[0] = grp->exchanges;
    out[1] = grp->bytes_moved;
    return B200GEO_OK;
}

}
