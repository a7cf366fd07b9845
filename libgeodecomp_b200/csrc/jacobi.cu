// Jacobi 6/7/27-point sweeps on one f64 member (B200GEO_KERNEL_JACOBI6/7/27).
//
// Replaces, for the bound models, the reference's Cell::update / Cell::updateLineX inner loop
// reached through FixedNeighborhoodUpdateFunctor / LinePointerUpdateFunctor
// (storage/fixedneighborhoodupdatefunctor.h:126-256, storage/linepointerupdatefunctor.h:56-180).
// The floating-point expression trees are those of oracle/models/jacobi.h, term by term and in
// the same association order, and contain no a*b+c site, so results are bit-identical to the
// reference's SerialSimulator.
//
// Roofline: HBM-bound, 16 algorithmic bytes per lattice update (one f64 read + one f64 write).
// Each thread owns two x-adjacent cells (one aligned 128-bit access) and marches along z, so a
// cell's z-neighbours (7-point) or whole plane sums (27-point) stay in registers; x/y neighbours
// are re-read through L1/L2, never from HBM (three planes of a 1024^2 slab are 25 MB << 126 MB L2).
#include "grid.h"

#include <cstring>

namespace b200geo {

namespace {

__device__ __forceinline__ double2 ld2(const double *p)
{
    return *reinterpret_cast<const double2 *>(p);
}

// pull a future plane's line from HBM into L2 without tying up a register or a scoreboard slot
__device__ __forceinline__ void prefetch_l2(const double *p)
{
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// (W + C) + E for the two cells of a pair in row p (p points at the pair's first cell)
__device__ __forceinline__ double2 row_sum(const double *p)
{
    double2 c = ld2(p);
    double w = p[-1], e = p[2];
    double2 r;
    r.x = (w + c.x) + c.y;
    r.y = (c.x + c.y) + e;
    return r;
}

__device__ __forceinline__ double2 plane_sum(const double *p, int64_t pitch)
{
    double2 a = row_sum(p - pitch), b = row_sum(p), c = row_sum(p + pitch);
    double2 r;
    r.x = (a.x + b.x) + c.x;
    r.y = (a.y + b.y) + c.y;
    return r;
}

template<int KIND>
__global__ void __launch_bounds__(256)
jacobi_kernel(const double *__restrict__ src, double *__restrict__ dst, int64_t pitch, int64_t plane,
              Box box, int xa, int zchunk, int pf, int zlast, int pdl)
{
    // programmatic dependent launch (small, L2-resident grids): let the NEXT sweep's CTAs be scheduled while
    // this one drains, and do not touch memory before the previous sweep has completed and flushed
    if (pdl) {
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        asm volatile("griddepcontrol.wait;" ::: "memory");
    }
    const int x = xa + 2 * (blockIdx.x * blockDim.x + threadIdx.x);
    const int y = box.y0 + blockIdx.y * blockDim.y + threadIdx.y;
    const int zb = box.z0 + blockIdx.z * zchunk;
    if (x >= box.x1 || y >= box.y1) return;
    const int ze = min(zb + zchunk, box.z1);
    const bool v0 = x >= box.x0, v1 = x + 1 < box.x1;

    const double *p = src + (int64_t)zb * plane + (int64_t)y * pitch + x;
    double *q = dst + (int64_t)zb * plane + (int64_t)y * pitch + x;

    if (KIND == 27) {
        double2 pm = plane_sum(p - plane, pitch), pc = plane_sum(p, pitch);
        for (int z = zb; z < ze; ++z, p += plane, q += plane) {
            if (pf && z + pf <= zlast) prefetch_l2(p + (int64_t)pf * plane);
            double2 pp = plane_sum(p + plane, pitch);
            double2 r;
            r.x = ((pm.x + pc.x) + pp.x) * (1.0 / 27.0);
            r.y = ((pm.y + pc.y) + pp.y) * (1.0 / 27.0);
            if (v0 && v1) *reinterpret_cast<double2 *>(q) = r;
            else if (v0) q[0] = r.x;
            else if (v1) q[1] = r.y;
            pm = pc;
            pc = pp;
        }
    } else {
        double2 zm = ld2(p - plane), c = ld2(p);
        for (int z = zb; z < ze; ++z, p += plane, q += plane) {
            if (pf && z + pf <= zlast) prefetch_l2(p + (int64_t)pf * plane);
            double2 zp = ld2(p + plane);
            double2 ym = ld2(p - pitch), yp = ld2(p + pitch);
            double w = p[-1], e = p[2];
            double2 r;
            if (KIND == 6) {
                r.x = (zm.x + ym.x + w + c.y + yp.x + zp.x) * (1.0 / 6.0);
                r.y = (zm.y + ym.y + c.x + e + yp.y + zp.y) * (1.0 / 6.0);
            } else {
                r.x = (zm.x + ym.x + w + c.x + c.y + yp.x + zp.x) * (1.0 / 7.0);
                r.y = (zm.y + ym.y + c.x + c.y + e + yp.y + zp.y) * (1.0 / 7.0);
            }
            if (v0 && v1) *reinterpret_cast<double2 *>(q) = r;
            else if (v0) q[0] = r.x;
            else if (v1) q[1] = r.y;
            zm = c;
            c = zp;
        }
    }
}

}

int sweep_jacobi(b200geo_grid *g, int kind, const Box& box, cudaStream_t s)
{
    const MemberLayout& L = g->m[0];
    const double *src = (const double *)g->member_ptr(0, 0) + L.origin;
    double *dst = (double *)g->member_ptr(0, 1) + L.origin;
    int xa = box.x0 & ~1;
    int pairs = (box.x1 - xa + 1) / 2;
    int ny = box.y1 - box.y0, nz = box.z1 - box.z0;
    dim3 block(64, 4);
    if (pairs <= 32) block = dim3(32, 8);
    int gx = (pairs + block.x - 1) / block.x, gy = (ny + block.y - 1) / block.y;
    // enough z chunks for >= 8 CTAs per SM, but long enough to amortise the two-plane prologue (measured on
    // 64^3 .. 256^3 grids, profiles/r1s_tuning.md: 2 planes for the 6/7-point kernels, 4 for the 27-point one)
    int zchunk = 32;
    const int zmin = kind == 27 ? 4 : 2;
    while (zchunk > zmin && (int64_t)gx * gy * ((nz + zchunk - 1) / zchunk) < 148 * 8) zchunk /= 2;
    if (g_tuning.jacobi_zchunk > 0) zchunk = g_tuning.jacobi_zchunk;
    // measured on B200 (profiles/r1b_tuning.md): +14% for the 7-point kernel, nothing for the 27-point one
    int pf = g_tuning.jacobi_prefetch >= 0 ? g_tuning.jacobi_prefetch : (kind == 27 ? 0 : 2), zlast = g->d[2] + g->g[2] - 1;
    dim3 grid(gx, gy, (nz + zchunk - 1) / zchunk);
    if (grid.y > 65535 || grid.z > 65535) return fail(B200GEO_ERR_OUT_OF_RANGE, "grid dimension too large");
    // A sweep over a small grid (both buffers L2 resident) costs about as much as the gap between two
    // dependent launches: overlap the launch of sweep n + 1 with the tail of sweep n (programmatic stream
    // serialization). Large grids gain nothing and keep the plain launch.
    const int64_t cells = (int64_t)(box.x1 - box.x0) * ny * nz;
    const int pdl = g_tuning.jacobi_pdl >= 0 ? g_tuning.jacobi_pdl : (cells <= (1 << 24) ? 1 : 0);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    const int64_t pitch = L.pitch, plane = L.plane;
    cudaError_t e;
    switch (kind) {
    case 6: e = cudaLaunchKernelEx(&cfg, jacobi_kernel<6>, src, dst, pitch, plane, box, xa, zchunk, pf, zlast, pdl); break;
    case 7: e = cudaLaunchKernelEx(&cfg, jacobi_kernel<7>, src, dst, pitch, plane, box, xa, zchunk, pf, zlast, pdl); break;
    default: e = cudaLaunchKernelEx(&cfg, jacobi_kernel<27>, src, dst, pitch, plane, box, xa, zchunk, pf, zlast, pdl); break;
    }
    count_launch();
    return check_cuda(e != cudaSuccess ? e : cudaGetLastError(), "jacobi sweep");
}

}
