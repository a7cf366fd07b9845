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

int tensor_map(b200geo_grid *g, int which, int tile_cols, int tile_rows, CUtensorMap *out);

namespace {

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

// CX x-adjacent cells of one row, owned by one lane (CX = 2 or 4; 16-byte aligned in memory)
template<int CX>
struct Cells {
    double v[CX];
};

template<int CX>
__device__ __forceinline__ Cells<CX> zero_cells()
{
    Cells<CX> c;
#pragma unroll
    for (int k = 0; k < CX; ++k) c.v[k] = 0.0;
    return c;
}

template<int CX>
__device__ __forceinline__ Cells<CX> load_cells(const double *p)
{
    Cells<CX> c;
#pragma unroll
    for (int k = 0; k < CX; k += 2) {
        double2 d = *reinterpret_cast<const double2 *>(p + k);
        c.v[k] = d.x;
        c.v[k + 1] = d.y;
    }
    return c;
}

template<int CX>
__device__ __forceinline__ void store_cells(double *p, const Cells<CX>& c)
{
#pragma unroll
    for (int k = 0; k < CX; k += 2) *reinterpret_cast<double2 *>(p + k) = make_double2(c.v[k], c.v[k + 1]);
}

// (W + C) + E for the cells of a lane (27-point row sums, same association as jacobi.cu): the lane's outer
// neighbours come from the adjacent lanes by shuffle — two shuffles per CX cells
template<int CX>
__device__ __forceinline__ Cells<CX> row_sum(const Cells<CX>& c)
{
    const double w = shfl_up1(c.v[CX - 1]), e = shfl_down1(c.v[0]);
    Cells<CX> r;
#pragma unroll
    for (int k = 0; k < CX; ++k) {
        const double l = k == 0 ? w : c.v[k == 0 ? 0 : k - 1];
        const double h = k == CX - 1 ? e : c.v[k == CX - 1 ? k : k + 1];
        r.v[k] = (l + c.v[k]) + h;
    }
    return r;
}

// Per-level pipeline state of one thread: R rows x CX cells.
//  6/7-point: zm = plane p-1, acc = partial sum of plane p-1's update still waiting for plane p
//  27-point : zm = plane sum S(p-2), acc = plane sum S(p-1)
template<int CX, int R>
struct LevelState {
    Cells<CX> zm[R], acc[R];
};

// Feed plane p of one level (own rows n[], the row above and the row below) and get the next
// level's plane p-1.
template<int KIND, int CX, int R>
__device__ __forceinline__ void feed_plane(LevelState<CX, R>& st, const Cells<CX> (&n)[R], const Cells<CX>& up, const Cells<CX>& dn,
                                           Cells<CX> (&out)[R])
{
    if (KIND == 27) {
        Cells<CX> rs[R + 2];
        rs[0] = row_sum<CX>(up);
#pragma unroll
        for (int r = 0; r < R; ++r) rs[r + 1] = row_sum<CX>(n[r]);
        rs[R + 1] = row_sum<CX>(dn);
#pragma unroll
        for (int r = 0; r < R; ++r) {
#pragma unroll
            for (int k = 0; k < CX; ++k) {
                const double s = (rs[r].v[k] + rs[r + 1].v[k]) + rs[r + 2].v[k];
                out[r].v[k] = ((st.zm[r].v[k] + st.acc[r].v[k]) + s) * (1.0 / 27.0);
                st.zm[r].v[k] = st.acc[r].v[k];
                st.acc[r].v[k] = s;
            }
        }
    } else {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const Cells<CX> c = n[r];
            const Cells<CX>& ym = r == 0 ? up : n[r == 0 ? 0 : r - 1];
            const Cells<CX>& yp = r == R - 1 ? dn : n[r == R - 1 ? r : r + 1];
            const double w = shfl_up1(c.v[CX - 1]), e = shfl_down1(c.v[0]);
#pragma unroll
            for (int k = 0; k < CX; ++k) {
                const double l = k == 0 ? w : c.v[k == 0 ? 0 : k - 1];
                const double h = k == CX - 1 ? e : c.v[k == CX - 1 ? k : k + 1];
                if (KIND == 6) {
                    out[r].v[k] = (st.acc[r].v[k] + c.v[k]) * (1.0 / 6.0);
                    st.acc[r].v[k] = st.zm[r].v[k] + ym.v[k] + l + h + yp.v[k];
                } else {
                    out[r].v[k] = (st.acc[r].v[k] + c.v[k]) * (1.0 / 7.0);
                    st.acc[r].v[k] = st.zm[r].v[k] + ym.v[k] + l + c.v[k] + h + yp.v[k];
                }
            }
            st.zm[r] = c;
        }
    }
}

// cells [first, first + CX) of a lane: those outside [lo, hi) take the edge value (o = the whole row is outside)
template<int CX>
__device__ __forceinline__ void with_edge(Cells<CX>& c, bool o, unsigned ox, double edge)
{
#pragma unroll
    for (int k = 0; k < CX; ++k)
        if (o || ((ox >> k) & 1)) c.v[k] = edge;
}

// KIND: 6 / 7 / 27 points; T sweeps per launch; a lane owns CX x-adjacent cells of R rows, a warp R rows of
// TX = 32 * CX cells, a CTA NW warps = a TX x (R * NW) tile; NS shared-memory stages; H = halo in x (>= T, even:
// 128-bit stores stay aligned; a multiple of 4 keeps the stored core a whole number of 64-byte blocks)
template<int KIND, int T, int CX, int R, int NW, int NS, int MINB, int H>
__global__ void __launch_bounds__(NW * 32, MINB)
jacobi_tb_kernel(const __grid_constant__ CUtensorMap tmap, double *__restrict__ dst, int64_t pitch, int64_t plane,
                 Box box, int xa, Limits lim, double edge, int zchunk, int pad_x, int pad_y, int pad_z, int gx, int gy, int panel)
{
    static_assert(CX == 2 || CX == 4, "a lane owns 2 or 4 cells of a row");
    static_assert(H >= T && H % 2 == 0, "x halo");
    constexpr int TX = 32 * CX;
    constexpr int TY = R * NW;
    constexpr int NX = T > 1 ? T - 1 : 1;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *stage = reinterpret_cast<double *>(smem_raw);  // [NS][TY][TX]
    double *xch = stage + NS * TY * TX;                    // [NX][2][NW][2][TX]: first and last row of every warp
    uint64_t *bars = reinterpret_cast<uint64_t *>(xch + NX * 2 * NW * 2 * TX);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // CTA order: tiles are numbered along x first (panel = 0), or in panels of `panel` tile rows with y fastest
    // inside a panel — neighbouring CTAs then share their longer (x) halos as well as the y ones
    int bx, by;
    if (panel <= 0) {
        bx = blockIdx.x % gx;
        by = blockIdx.x / gx;
    } else {
        const int per_panel = panel * gx, p = blockIdx.x / per_panel, rest = blockIdx.x % per_panel;
        const int rows = min(panel, gy - p * panel);
        bx = rest / rows;
        by = p * panel + rest % rows;
    }
    const int X0 = xa + bx * (TX - 2 * H) - H;
    const int Y0 = box.y0 + by * (TY - 2 * T) - T;
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

    // this thread's cells: x .. x + CX - 1 in rows Y0 + warp * R + r
    const int x = X0 + CX * lane;
    const int yr = Y0 + warp * R;
    // bit k of ox: cell x + k lies outside the simulation area; bit k of sx: cell x + k is stored
    unsigned ox = 0, sx = 0;
#pragma unroll
    for (int k = 0; k < CX; ++k) {
        if (x + k < lim.lo[0] || x + k >= lim.hi[0]) ox |= 1u << k;
        if (CX * lane + k >= H && CX * lane + k < TX - H && x + k >= box.x0 && x + k < box.x1) sx |= 1u << k;
    }
    // bit r + 1 of oy: row yr + r lies outside the simulation area (r = -1 .. R)
    unsigned oy = 0;
#pragma unroll
    for (int r = -1; r <= R; ++r)
        if (yr + r < lim.lo[1] || yr + r >= lim.hi[1]) oy |= 1u << (r + 1);
    const bool sall = sx == (1u << CX) - 1;
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

    LevelState<CX, R> st[T];
    Cells<CX> carry[T][R];  // carry[t]: level t plane produced in the previous iteration (t >= 1)
#pragma unroll
    for (int t = 0; t < T; ++t)
#pragma unroll
        for (int r = 0; r < R; ++r) {
            st[t].zm[r] = zero_cells<CX>();
            st[t].acc[r] = zero_cells<CX>();
            carry[t][r] = zero_cells<CX>();
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
            Cells<CX> n[R], up, dn, out[R];
            if (t == 0) {
                if (i >= nload) continue;
                const int s = i % NS;
                while (!mbar_try_wait(&bars[s], (i / NS) & 1)) {}
                const double *sp = stage + s * TY * TX + CX * lane;
                const int z = zs + i;
                const bool oz = z < lim.lo[2] || z >= lim.hi[2];
#pragma unroll
                for (int r = 0; r < R; ++r) n[r] = load_cells<CX>(sp + (warp * R + r) * TX);
                up = load_cells<CX>(sp + up_row * TX);
                dn = load_cells<CX>(sp + dn_row * TX);
                if (edge_xy || oz) {  // uniform branch: interior tiles never take it
#pragma unroll
                    for (int r = 0; r < R; ++r) with_edge<CX>(n[r], oz || ((oy >> (r + 1)) & 1), ox, edge);
                    with_edge<CX>(up, oz || (oy & 1), ox, edge);
                    with_edge<CX>(dn, oz || ((oy >> (R + 1)) & 1), ox, edge);
                }
            } else {
                const double *xp = xch + (((t - 1) * 2 + (par ^ 1)) * NW) * 2 * TX + CX * lane;
#pragma unroll
                for (int r = 0; r < R; ++r) n[r] = carry[t][r];
                up = load_cells<CX>(xp + (up_warp * 2 + 1) * TX);
                dn = load_cells<CX>(xp + (dn_warp * 2 + 0) * TX);
            }
            feed_plane<KIND, CX, R>(st[t], n, up, dn, out);
            const int zo = zs + i - 2 * t - 1;  // plane index of `out`, a plane of level t + 1
            const bool oz = zo < lim.lo[2] || zo >= lim.hi[2];
            if (edge_xy || oz) {
#pragma unroll
                for (int r = 0; r < R; ++r) with_edge<CX>(out[r], oz || ((oy >> (r + 1)) & 1), ox, edge);
            }
            if (t == T - 1) {
                if (zo >= zb && zo < ze) {
                    double *q = q0 + (int64_t)zo * plane;
                    if (sall) {  // the common case: predicated 128-bit stores, no branches per row
#pragma unroll
                        for (int r = 0; r < R; ++r)
                            if ((sy >> r) & 1) store_cells<CX>(q + r * pitch, out[r]);
                    } else if (sx) {  // tile halo, odd box edges in x
#pragma unroll
                        for (int r = 0; r < R; ++r)
                            if ((sy >> r) & 1) {
#pragma unroll
                                for (int k = 0; k < CX; ++k)
                                    if ((sx >> k) & 1) q[r * pitch + k] = out[r].v[k];
                            }
                    }
                }
            } else {
                double *xp = xch + ((t * 2 + par) * NW + warp) * 2 * TX + CX * lane;
                store_cells<CX>(xp, out[0]);
                store_cells<CX>(xp + TX, out[R - 1]);
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

template<int KIND, int T, int CX, int R, int NW, int NS, int MINB, int H>
int launch_tb(b200geo_grid *g, const Box& box, const Limits& lim, double edge, cudaStream_t s)
{
    constexpr int TX = 32 * CX;
    constexpr int TY = R * NW;
    constexpr int NX = T > 1 ? T - 1 : 1;
    const MemberLayout& L = g->m[0];
    CUtensorMap map;
    int rc = tensor_map(g, g->cur, TX, TY, &map);
    if (rc) return rc;
    size_t smem = (size_t)NS * TY * TX * 8 + (size_t)NX * 2 * NW * 2 * TX * 8 + NS * 8;
    auto kernel = jacobi_tb_kernel<KIND, T, CX, R, NW, NS, MINB, H>;
    // function attributes are per device: one process may drive several GPUs (slab groups)
    static bool attr_set[64] = {false};
    if (g->device < 0 || g->device >= 64 || !attr_set[g->device]) {
        B200GEO_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (g->device >= 0 && g->device < 64) attr_set[g->device] = true;
    }
    int xa = box.x0 & ~(CX - 1);
    int gx = (box.x1 - xa + (TX - 2 * H) - 1) / (TX - 2 * H);
    int gy = (box.y1 - box.y0 + (TY - 2 * T) - 1) / (TY - 2 * T);
    int nz = box.z1 - box.z0;
    // long z chunks amortise the 2T warm-up planes; enough chunks to fill the machine several times over
    // (deeper blocking warms up on more planes per chunk: 256 planes measured 1.3 % faster than 128 for T = 4, profiles/r3d)
    int zchunk = g_tuning.jacobi_tb_zchunk > 0 ? g_tuning.jacobi_tb_zchunk : (T >= 3 ? 256 : 128);
    while (zchunk > 16 && (int64_t)gx * gy * ((nz + zchunk - 1) / zchunk) < 148 * 6) zchunk /= 2;
    dim3 grid((unsigned)(gx * gy), 1, (nz + zchunk - 1) / zchunk);
    if ((int64_t)gx * gy > 0x7fffffff || grid.z > 65535) return fail(B200GEO_ERR_OUT_OF_RANGE, "grid dimension too large");
    double *dst = (double *)g->member_ptr(0, 1) + L.origin;
    kernel<<<grid, NW * 32, smem, s>>>(map, dst, L.pitch, L.plane, box, xa, lim, edge, zchunk, L.lead, g->g[1], g->g[2], gx, gy,
                                       g_tuning.jacobi_tb_raster);
    count_launch();
    return check_cuda(cudaGetLastError(), "temporal-blocked jacobi sweep");
}

template<int KIND, int T>
int launch_tb_shape(b200geo_grid *g, int rows, const Box& box, const Limits& lim, double edge, cudaStream_t s)
{
    // tile shapes ("jacobi.tb_rows"): the register budget (state = 6 f64 per cell and level) decides how many CTAs
    // share an SM, the halo (2H of TX columns, 2T of TY rows) how much of a tile is redundant
    constexpr int HD = (T + 1) & ~1;        // narrowest even x halo
    constexpr int H4 = (T + 3) & ~3;        // x halo that keeps stored cores whole multiples of 64 bytes
    constexpr int M2 = T == 2 ? 2 : 1;
    switch (rows) {
    case 64: return launch_tb<KIND, T, 2, 4, 16, 3, 1, HD>(g, box, lim, edge, s);    // 64 x 64, 4 rows / thread
    case 34: return launch_tb<KIND, T, 2, 4, 8, 4, 1, HD>(g, box, lim, edge, s);     // 64 x 32, one CTA / SM
    case 31: return launch_tb<KIND, T, 2, 2, 16, 4, 1, HD>(g, box, lim, edge, s);
    case 32: return launch_tb<KIND, T, 2, 2, 16, 4, M2, HD>(g, box, lim, edge, s);   // 64 x 32, 2 rows / thread
    case 40: return launch_tb<KIND, T, 2, 4, 8, 4, M2, H4>(g, box, lim, edge, s);    // as 33 with 64-byte-aligned cores
    default: return launch_tb<KIND, T, 2, 4, 8, 4, M2, HD>(g, box, lim, edge, s);    // 33: 64 x 32, 4 rows / thread
    }
}

template<int KIND>
int launch_tb_depth(b200geo_grid *g, int depth, int rows, const Box& box, const Limits& lim, double edge, cudaStream_t s)
{
    switch (depth) {
    case 2: return launch_tb_shape<KIND, 2>(g, rows, box, lim, edge, s);
    case 3: return launch_tb_shape<KIND, 3>(g, rows, box, lim, edge, s);
    case 4: return launch_tb_shape<KIND, 4>(g, rows, box, lim, edge, s);
    default: return fail(B200GEO_ERR_INVALID, "temporal blocking depth must be 2, 3 or 4");
    }
}

}

void *tensor_map_encoder()
{
    return (void *)encode_tiled();
}

// TMA descriptor of member 0 of buffer `which` (absolute index): the whole padded array as a
// rank-3 tensor of f64, box = one TX x TY x 1 tile.
int tensor_map(b200geo_grid *g, int which, int tile_cols, int tile_rows, CUtensorMap *out)
{
    EncodeTiled enc = encode_tiled();
    if (!enc) return fail(B200GEO_ERR_CUDA, "CUDA error: cuTensorMapEncodeTiled is not available in this driver");
    const MemberLayout& L = g->m[0];
    cuuint64_t dims[3] = {(cuuint64_t)L.pitch, (cuuint64_t)(g->d[1] + 2 * g->g[1]), (cuuint64_t)(g->d[2] + 2 * g->g[2])};
    cuuint64_t strides[2] = {(cuuint64_t)L.pitch * 8, (cuuint64_t)L.plane * 8};
    cuuint32_t boxdim[3] = {(cuuint32_t)tile_cols, (cuuint32_t)tile_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    const CUtensorMapL2promotion promo = g_tuning.jacobi_tb_promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE :
        g_tuning.jacobi_tb_promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B :
        g_tuning.jacobi_tb_promo == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, g->buf[which] + L.offset, dims, strides, boxdim, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, promo,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(B200GEO_ERR_CUDA, "CUDA error: cuTensorMapEncodeTiled failed");
    return B200GEO_OK;
}

int sweep_jacobi_tb(b200geo_grid *g, int kind, int depth, const Box& box, cudaStream_t s)
{
    int rows = g_tuning.jacobi_tb_rows;
    Limits lim;
    for (int i = 0; i < 3; ++i) {
        lim.lo[i] = g->desc.ghost_mode[i][0] == B200GEO_GHOST_EDGE ? 0 : INT_MIN;
        lim.hi[i] = g->desc.ghost_mode[i][1] == B200GEO_GHOST_EDGE ? g->d[i] : INT_MAX;
    }
    double edge;
    memcpy(&edge, g->edge + g->m[0].edge_offset, 8);
    switch (kind) {
    case 6: return launch_tb_depth<6>(g, depth, rows, box, lim, edge, s);
    case 7: return launch_tb_depth<7>(g, depth, rows, box, lim, edge, s);
    default: return launch_tb_depth<27>(g, depth, rows, box, lim, edge, s);
    }
}

}
