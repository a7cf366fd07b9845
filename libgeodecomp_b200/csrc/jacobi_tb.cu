// Temporal-blocked Jacobi 6/7/27-point kernels: T sweeps per HBM round trip.
//
// Same arithmetic per cell and sweep as jacobi.cu (and therefore as the reference's
// Cell::update / updateLineX, see there), so T launches of the single-sweep kernel and one launch
// of this one give bit-identical grids. What changes is the traffic: a cell is read from HBM once
// and written once per T sweeps, i.e. 16/T algorithmic bytes per lattice update.
//
// Scheme ("3.5-D blocking"): a CTA owns a 64 x TY column of cells and streams along z.
//  - level 0 (the input grid) arrives plane by plane through TMA (cp.async.bulk.tensor.3d, one
//    64 x TY x 1 box per plane) into an NS-deep ring of shared-memory stages guarded by
//    mbarriers; out-of-array parts of a box are zero-filled by the TMA unit, cells outside the
//    simulation area on a Cube axis are replaced by the constant edge cell at EVERY level
//    (the reference's padding ring never changes, storage/soagrid.h:578-584).
//  - a warp owns R whole rows of 64 cells, a lane two x-adjacent cells of each: x neighbours come
//    from warp shuffles, y neighbours inside the R rows from registers, the two rows above and
//    below from a small shared-memory exchange buffer, z neighbours from registers (each level
//    keeps a running partial sum of the planes it has seen).
//  - levels are skewed by two planes so that ONE __syncthreads per streamed plane is enough: in
//    iteration i level t consumes what level t-1 produced in iteration i-1.
//  - the outermost ring of each level is garbage (no neighbours); after T levels the valid core is
//    (64 - 2H) x (TY - 2T) cells (H = T rounded up to even, so that 128-bit stores stay aligned),
//    which is what the CTA stores. Neighbouring CTAs overlap by the halo; those re-reads hit L2.
#include "grid.h"

#include <cuda.h>

#include <climits>
#include <cstring>

namespace b200geo {

namespace {

constexpr int TX = 64;

struct Limits {
    int lo[3], hi[3];  // cells outside [lo, hi) on an axis hold the edge cell (Cube sides only)
};

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

__device__ __forceinline__ void tma_load_plane(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

__device__ __forceinline__ double shfl_up1(double v)
{
    return __shfl_up_sync(0xffffffffu, v, 1);
}

__device__ __forceinline__ double shfl_down1(double v)
{
    return __shfl_down_sync(0xffffffffu, v, 1);
}

// (W + C) + E for the two cells of a lane (27-point row sums, same association as jacobi.cu)
__device__ __forceinline__ double2 row_sum(double2 c)
{
    double w = shfl_up1(c.y), e = shfl_down1(c.x);
    double2 r;
    r.x = (w + c.x) + c.y;
    r.y = (c.x + c.y) + e;
    return r;
}

// Per-level pipeline state of one thread: R rows x 2 cells.
//  6/7-point: zm = plane p-1, acc = partial sum of plane p-1's update still waiting for plane p
//  27-point : zm = plane sum S(p-2), acc = plane sum S(p-1)
template<int R>
struct LevelState {
    double2 zm[R], acc[R];
};

// Feed plane p of one level (own rows n[], the row above and the row below) and get the next
// level's plane p-1.
template<int KIND, int R>
__device__ __forceinline__ void feed_plane(LevelState<R>& st, const double2 (&n)[R], double2 up, double2 dn, double2 (&out)[R])
{
    if (KIND == 27) {
        double2 rs[R + 2];
        rs[0] = row_sum(up);
#pragma unroll
        for (int r = 0; r < R; ++r) rs[r + 1] = row_sum(n[r]);
        rs[R + 1] = row_sum(dn);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            double2 s;
            s.x = (rs[r].x + rs[r + 1].x) + rs[r + 2].x;
            s.y = (rs[r].y + rs[r + 1].y) + rs[r + 2].y;
            out[r].x = ((st.zm[r].x + st.acc[r].x) + s.x) * (1.0 / 27.0);
            out[r].y = ((st.zm[r].y + st.acc[r].y) + s.y) * (1.0 / 27.0);
            st.zm[r] = st.acc[r];
            st.acc[r] = s;
        }
    } else {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            double2 c = n[r];
            double2 ym = r == 0 ? up : n[r == 0 ? 0 : r - 1];
            double2 yp = r == R - 1 ? dn : n[r == R - 1 ? r : r + 1];
            double w = shfl_up1(c.y), e = shfl_down1(c.x);
            if (KIND == 6) {
                out[r].x = (st.acc[r].x + c.x) * (1.0 / 6.0);
                out[r].y = (st.acc[r].y + c.y) * (1.0 / 6.0);
                st.acc[r].x = st.zm[r].x + ym.x + w + c.y + yp.x;
                st.acc[r].y = st.zm[r].y + ym.y + c.x + e + yp.y;
            } else {
                out[r].x = (st.acc[r].x + c.x) * (1.0 / 7.0);
                out[r].y = (st.acc[r].y + c.y) * (1.0 / 7.0);
                st.acc[r].x = st.zm[r].x + ym.x + w + c.x + c.y + yp.x;
                st.acc[r].y = st.zm[r].y + ym.y + c.x + c.y + e + yp.y;
            }
            st.zm[r] = c;
        }
    }
}

__device__ __forceinline__ double2 with_edge(double2 v, bool o0, bool o1, double edge)
{
    if (o0) v.x = edge;
    if (o1) v.y = edge;
    return v;
}

template<int KIND, int T, int R, int NW, int NS, int MINB>
__global__ void __launch_bounds__(NW * 32, MINB)
jacobi_tb_kernel(const __grid_constant__ CUtensorMap tmap, double *__restrict__ dst, int64_t pitch, int64_t plane,
                 Box box, int xa, Limits lim, double edge, int zchunk, int pad_x, int pad_y, int pad_z)
{
    constexpr int TY = R * NW;
    constexpr int H = (T + 1) & ~1;
    constexpr int NX = T > 1 ? T - 1 : 1;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *stage = reinterpret_cast<double *>(smem_raw);  // [NS][TY][TX]
    double *xch = stage + NS * TY * TX;                    // [NX][2][NW][2][TX]: first and last row of every warp
    uint64_t *bars = reinterpret_cast<uint64_t *>(xch + NX * 2 * NW * 2 * TX);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int X0 = xa + blockIdx.x * (TX - 2 * H) - H;
    const int Y0 = box.y0 + blockIdx.y * (TY - 2 * T) - T;
    const int zb = box.z0 + blockIdx.z * zchunk;
    const int ze = min(zb + zchunk, box.z1);
    const int zs = zb - T;
    const int nload = (ze - zb) + 2 * T;
    const int niter = (ze - zb) + 3 * T - 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int s = 0; s < NS - 1 && s < nload; ++s) {
            mbar_expect_tx(&bars[s], TY * TX * 8);
            tma_load_plane(stage + s * TY * TX, &tmap, &bars[s], X0 + pad_x, Y0 + pad_y, zs + s + pad_z);
        }
    }

    // this thread's cells: x, x + 1 in rows Y0 + warp * R + r
    const int x = X0 + 2 * lane;
    const int yr = Y0 + warp * R;
    const bool ox0 = x < lim.lo[0] || x >= lim.hi[0], ox1 = x + 1 < lim.lo[0] || x + 1 >= lim.hi[0];
    // bit r + 1 of oy: row yr + r lies outside the simulation area (r = -1 .. R)
    unsigned oy = 0;
#pragma unroll
    for (int r = -1; r <= R; ++r)
        if (yr + r < lim.lo[1] || yr + r >= lim.hi[1]) oy |= 1u << (r + 1);
    const bool sx0 = 2 * lane >= H && 2 * lane < TX - H && x >= box.x0 && x < box.x1;
    const bool sx1 = 2 * lane >= H && 2 * lane < TX - H && x + 1 >= box.x0 && x + 1 < box.x1;
    const bool sboth = sx0 && sx1;
    unsigned sy = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        int ty = warp * R + r;
        if (ty >= T && ty < TY - T && yr + r >= box.y0 && yr + r < box.y1) sy |= 1u << r;
    }
    // CTA-uniform: does this tile touch cells outside the simulation area in x or y at all?
    const bool edge_xy = X0 < lim.lo[0] || X0 + TX > lim.hi[0] || Y0 < lim.lo[1] || Y0 + TY > lim.hi[1];
    const int up_row = warp * R > 0 ? warp * R - 1 : 0;
    const int dn_row = warp * R + R < TY ? warp * R + R : TY - 1;
    const int up_warp = warp > 0 ? warp - 1 : 0, dn_warp = warp + 1 < NW ? warp + 1 : NW - 1;

    LevelState<R> st[T];
    double2 carry[T][R];  // carry[t]: level t plane produced in the previous iteration (t >= 1)
#pragma unroll
    for (int t = 0; t < T; ++t)
#pragma unroll
        for (int r = 0; r < R; ++r) {
            st[t].zm[r] = make_double2(0.0, 0.0);
            st[t].acc[r] = make_double2(0.0, 0.0);
            carry[t][r] = make_double2(0.0, 0.0);
        }

    double *q0 = dst + (int64_t)yr * pitch + x;

#pragma unroll 2
    for (int i = 0; i < niter; ++i) {
        // refill the stage consumed in the previous iteration
        if (threadIdx.x == 0 && i + NS - 1 < nload) {
            int s = (i + NS - 1) % NS;
            mbar_expect_tx(&bars[s], TY * TX * 8);
            tma_load_plane(stage + s * TY * TX, &tmap, &bars[s], X0 + pad_x, Y0 + pad_y, zs + i + NS - 1 + pad_z);
        }
        const int par = i & 1;
#pragma unroll
        for (int t = T - 1; t >= 0; --t) {
            double2 n[R], up, dn, out[R];
            if (t == 0) {
                if (i >= nload) continue;
                const int s = i % NS;
                while (!mbar_try_wait(&bars[s], (i / NS) & 1)) {}
                const double *sp = stage + s * TY * TX + 2 * lane;
                const int z = zs + i;
                const bool oz = z < lim.lo[2] || z >= lim.hi[2];
#pragma unroll
                for (int r = 0; r < R; ++r) n[r] = *reinterpret_cast<const double2 *>(sp + (warp * R + r) * TX);
                up = *reinterpret_cast<const double2 *>(sp + up_row * TX);
                dn = *reinterpret_cast<const double2 *>(sp + dn_row * TX);
                if (edge_xy || oz) {  // uniform branch: interior tiles never take it
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        bool o = oz || ((oy >> (r + 1)) & 1);
                        n[r] = with_edge(n[r], o || ox0, o || ox1, edge);
                    }
                    bool o = oz || (oy & 1);
                    up = with_edge(up, o || ox0, o || ox1, edge);
                    o = oz || ((oy >> (R + 1)) & 1);
                    dn = with_edge(dn, o || ox0, o || ox1, edge);
                }
            } else {
                const double *xp = xch + (((t - 1) * 2 + (par ^ 1)) * NW) * 2 * TX + 2 * lane;
#pragma unroll
                for (int r = 0; r < R; ++r) n[r] = carry[t][r];
                up = *reinterpret_cast<const double2 *>(xp + (up_warp * 2 + 1) * TX);
                dn = *reinterpret_cast<const double2 *>(xp + (dn_warp * 2 + 0) * TX);
            }
            feed_plane<KIND, R>(st[t], n, up, dn, out);
            const int zo = zs + i - 2 * t - 1;  // plane index of `out`, a plane of level t + 1
            const bool oz = zo < lim.lo[2] || zo >= lim.hi[2];
            if (edge_xy || oz) {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    bool o = oz || ((oy >> (r + 1)) & 1);
                    out[r] = with_edge(out[r], o || ox0, o || ox1, edge);
                }
            }
            if (t == T - 1) {
                if (zo >= zb && zo < ze) {
                    double *q = q0 + (int64_t)zo * plane;
                    if (sboth) {  // the common case: predicated 128-bit stores, no branches per row
#pragma unroll
                        for (int r = 0; r < R; ++r)
                            if ((sy >> r) & 1) *reinterpret_cast<double2 *>(q + r * pitch) = out[r];
                    } else if (sx0 || sx1) {  // odd box edges in x
#pragma unroll
                        for (int r = 0; r < R; ++r)
                            if ((sy >> r) & 1) q[r * pitch + (sx0 ? 0 : 1)] = sx0 ? out[r].x : out[r].y;
                    }
                }
            } else {
                double *xp = xch + ((t * 2 + par) * NW + warp) * 2 * TX + 2 * lane;
                *reinterpret_cast<double2 *>(xp) = out[0];
                *reinterpret_cast<double2 *>(xp + TX) = out[R - 1];
#pragma unroll
                for (int r = 0; r < R; ++r) carry[t + 1][r] = out[r];
            }
        }
        __syncthreads();
    }
}

typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiled encode_tiled()
{
    static EncodeTiled fn = 0;
    if (!fn) {
        void *p = 0;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiled)p;
    }
    return fn;
}

template<int KIND, int T, int R, int NW, int NS, int MINB>
int launch_tb(b200geo_grid *g, const CUtensorMap& map, const Box& box, const Limits& lim, double edge, cudaStream_t s)
{
    constexpr int TY = R * NW;
    constexpr int H = (T + 1) & ~1;
    constexpr int NX = T > 1 ? T - 1 : 1;
    const MemberLayout& L = g->m[0];
    size_t smem = (size_t)NS * TY * TX * 8 + (size_t)NX * 2 * NW * 2 * TX * 8 + NS * 8;
    auto kernel = jacobi_tb_kernel<KIND, T, R, NW, NS, MINB>;
    // function attributes are per device: one process may drive several GPUs (slab groups)
    static bool attr_set[64] = {false};
    if (g->device < 0 || g->device >= 64 || !attr_set[g->device]) {
        B200GEO_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (g->device >= 0 && g->device < 64) attr_set[g->device] = true;
    }
    int xa = box.x0 & ~1;
    int gx = (box.x1 - xa + (TX - 2 * H) - 1) / (TX - 2 * H);
    int gy = (box.y1 - box.y0 + (TY - 2 * T) - 1) / (TY - 2 * T);
    int nz = box.z1 - box.z0;
    // long z chunks amortise the 2T warm-up planes; enough chunks to fill the machine several times over
    int zchunk = g_tuning.jacobi_tb_zchunk > 0 ? g_tuning.jacobi_tb_zchunk : 128;
    while (zchunk > 16 && (int64_t)gx * gy * ((nz + zchunk - 1) / zchunk) < 148 * 6) zchunk /= 2;
    dim3 grid(gx, gy, (nz + zchunk - 1) / zchunk);
    if (grid.y > 65535 || grid.z > 65535) return fail(B200GEO_ERR_OUT_OF_RANGE, "grid dimension too large");
    double *dst = (double *)g->member_ptr(0, 1) + L.origin;
    kernel<<<grid, NW * 32, smem, s>>>(map, dst, L.pitch, L.plane, box, xa, lim, edge, zchunk, L.lead, g->g[1], g->g[2]);
    count_launch();
    return check_cuda(cudaGetLastError(), "temporal-blocked jacobi sweep");
}

template<int KIND, int T>
int launch_tb_shape(b200geo_grid *g, const CUtensorMap& map, int rows, const Box& box, const Limits& lim, double edge, cudaStream_t s)
{
    // tile shapes: 64 x 32 (R = 2, 16 warps), 64 x 64 (R = 4, 16 warps), 64 x 32 (R = 4, 8 warps);
    // the register budget (state = 4 f64 per cell and level) decides how many CTAs share an SM
    switch (rows) {
    case 64: return launch_tb<KIND, T, 4, 16, 3, 1>(g, map, box, lim, edge, s);
    case 33: return launch_tb<KIND, T, 4, 8, 4, (T == 2 ? 2 : 1)>(g, map, box, lim, edge, s);
    case 34: return launch_tb<KIND, T, 4, 8, 4, 1>(g, map, box, lim, edge, s);
    case 31: return launch_tb<KIND, T, 2, 16, 4, 1>(g, map, box, lim, edge, s);
    default: return launch_tb<KIND, T, 2, 16, 4, (T == 2 ? 2 : 1)>(g, map, box, lim, edge, s);
    }
}

template<int KIND>
int launch_tb_depth(b200geo_grid *g, const CUtensorMap& map, int depth, int rows, const Box& box, const Limits& lim, double edge, cudaStream_t s)
{
    switch (depth) {
    case 2: return launch_tb_shape<KIND, 2>(g, map, rows, box, lim, edge, s);
    case 3: return launch_tb_shape<KIND, 3>(g, map, rows, box, lim, edge, s);
    case 4: return launch_tb_shape<KIND, 4>(g, map, rows, box, lim, edge, s);
    default: return fail(B200GEO_ERR_INVALID, "temporal blocking depth must be 2, 3 or 4");
    }
}

}

// TMA descriptor of member 0 of buffer `which` (absolute index): the whole padded array as a
// rank-3 tensor of f64, box = one 64 x TY x 1 tile.
static int tensor_map(b200geo_grid *g, int which, int tile_rows, CUtensorMap *out)
{
    EncodeTiled enc = encode_tiled();
    if (!enc) return fail(B200GEO_ERR_CUDA, "CUDA error: cuTensorMapEncodeTiled is not available in this driver");
    const MemberLayout& L = g->m[0];
    cuuint64_t dims[3] = {(cuuint64_t)L.pitch, (cuuint64_t)(g->d[1] + 2 * g->g[1]), (cuuint64_t)(g->d[2] + 2 * g->g[2])};
    cuuint64_t strides[2] = {(cuuint64_t)L.pitch * 8, (cuuint64_t)L.plane * 8};
    cuuint32_t boxdim[3] = {(cuuint32_t)TX, (cuuint32_t)tile_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, g->buf[which] + L.offset, dims, strides, boxdim, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(B200GEO_ERR_CUDA, "CUDA error: cuTensorMapEncodeTiled failed");
    return B200GEO_OK;
}

int sweep_jacobi_tb(b200geo_grid *g, int kind, int depth, const Box& box, cudaStream_t s)
{
    int rows = g_tuning.jacobi_tb_rows;
    int tile_rows = rows == 64 ? 64 : 32;
    CUtensorMap map;
    int rc = tensor_map(g, g->cur, tile_rows, &map);
    if (rc) return rc;
    Limits lim;
    for (int i = 0; i < 3; ++i) {
        lim.lo[i] = g->desc.ghost_mode[i][0] == B200GEO_GHOST_EDGE ? 0 : INT_MIN;
        lim.hi[i] = g->desc.ghost_mode[i][1] == B200GEO_GHOST_EDGE ? g->d[i] : INT_MAX;
    }
    double edge;
    memcpy(&edge, g->edge + g->m[0].edge_offset, 8);
    switch (kind) {
    case 6: return launch_tb_depth<6>(g, map, depth, rows, box, lim, edge, s);
    case 7: return launch_tb_depth<7>(g, map, depth, rows, box, lim, edge, s);
    default: return launch_tb_depth<27>(g, map, depth, rows, box, lim, edge, s);
    }
}

}
