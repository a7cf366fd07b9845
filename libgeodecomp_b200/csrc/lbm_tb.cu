// D3Q19 BGK lattice-Boltzmann, TWO sweeps per HBM round trip (B200GEO_KERNEL_LBM_D3Q19, fused path).
//
// Same cell as lbm.cu (lbm_cell.h: the reference's src/examples/latticeboltzmann/main.cpp:62-229, term by term,
// -fmad=false), so one launch of this kernel and two launches of the one-sweep kernel give bit-identical grids.
// What changes is the traffic: 19 populations + the state are read from HBM once and 19 populations are written
// once per TWO lattice updates: 78 B per update instead of 156.
//
// Scheme: a CTA owns a column of 32 x (NR - 2) cells and streams along z.
//  - sweep 1 is computed for the column plus a ring of one cell (34 x NR cells per plane: warp w < NR owns the 32
//    aligned cells of row w, the last warp the 2 x NR cells left and right of them). Its input arrives by TMA
//    (cp.async.bulk.tensor.4d over the member-major grid, NST-deep mbarrier ring, issued NST planes ahead): per plane
//    20 windows of 40 x NR elements, one per population — EACH SHIFTED IN y AND z BY THE OFFSET ITS POPULATION IS PULLED
//    FROM (x offsets are compile-time displacements of the shared-memory read: a TMA window has to start on a multiple
//    of 16 bytes, tools/probe/tma_probe.cu), so that a thread finds all 19 values it pulls, and its state, in its own
//    row of the windows — no registers spent on loads in flight, no address arithmetic per population. The ring cells that neighbouring CTAs compute as well come
//    out of L2;
//  - its results — the time level nobody outside the CTA ever sees — live in shared memory only, in rings of planes
//    that are exactly as deep as their consumers need: a population that travels up (T, TW, TE, TN, TS) is pulled
//    two planes later (3 planes), one that stays in its plane or travels down is pulled by the plane computed in the
//    same iteration or read as a wall cell's own value one plane later (2 planes): 43 plane-populations, not 3 x 19;
//  - sweep 2 pulls from those rings — wall cells through the same accessor, whatever their state — and stores its
//    32-cell rows (whole 128-byte lines) to the scratch buffer;
//  - wall cells of sweep 1 find the populations they bounce back among the pulled ones (they are the same cells'
//    values) and read their own 19 from the grid (faces only: one L2 round trip for the CTAs that touch a face);
//  - cells outside the simulation area on a Cube axis (the reference's constant padding ring,
//    storage/soagrid.h:578-584) are the edge cell at the intermediate time level too.
#include "grid.h"
#include "lbm_cell.h"

#include <cuda.h>

#include <climits>
#include <cstring>
#include <cstdlib>

namespace b200geo {

void *tensor_map_encoder();  // cuTensorMapEncodeTiled (jacobi_tb.cu)

namespace {

using namespace lbm;

struct Limits {
    int lo[3], hi[3];  // cells outside [lo, hi) on an axis belong to the constant edge ring (Cube sides only)
};

struct EdgeCell {
    float f[19];
};

constexpr int ROW = 34;   // cells per row of the intermediate level: 32 + one on either side
constexpr int WROW = 40;  // elements per row of a TMA window: x0 - 4 .. x0 + 35 (a window starts on a multiple of 16 bytes)
constexpr int WLEAD = 4;  // elements of a window row before the tile's first column
constexpr int WINDOWS = 20;  // 19 populations + the state

// populations that travel up (pulled with Z = -1) live for three planes, all others for two
__host__ __device__ constexpr bool travels_up(int comp)
{
    return comp == T || comp == TW || comp == TE || comp == TN || comp == TS;
}

// first plane slot of a population's ring: [5 rings of 3 planes][14 rings of 2 planes]
__host__ __device__ constexpr int ring_base(int comp)
{
    return comp == T ? 0 : comp == TW ? 3 : comp == TE ? 6 : comp == TN ? 9 : comp == TS ? 12 :
           15 + 2 * (comp == C ? 0 : comp == N ? 1 : comp == E ? 2 : comp == W ? 3 : comp == S ? 4 : comp == B ? 5 :
                     comp == NW ? 6 : comp == SW ? 7 : comp == NE ? 8 : comp == SE ? 9 : comp == BW ? 10 : comp == BE ? 11 :
                     comp == BN ? 12 : 13);
}

constexpr int SLOTS = 43;

// where a population is pulled from (lbm_cell.h: pull()); the state and everything else: the cell itself
__host__ __device__ constexpr int pull_x(int comp)
{
    return comp == E || comp == NE || comp == SE || comp == TE || comp == BE ? -1 :
           comp == W || comp == NW || comp == SW || comp == TW || comp == BW ? 1 : 0;
}

__host__ __device__ constexpr int pull_y(int comp)
{
    return comp == N || comp == NW || comp == NE || comp == TN || comp == BN ? -1 :
           comp == S || comp == SW || comp == SE || comp == TS || comp == BS ? 1 : 0;
}

__host__ __device__ constexpr int pull_z(int comp)
{
    return travels_up(comp) ? -1 : comp == B || comp == BW || comp == BE || comp == BN || comp == BS ? 1 : 0;
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

__device__ __forceinline__ void tma_load_window(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2, int c3)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

__device__ __forceinline__ void tma_load_window_hint(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2, int c3,
                                                     uint64_t policy)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "l"(policy)
        : "memory");
}

__device__ __forceinline__ void store_out(float *p, float v, bool streaming)
{
    if (streaming) __stcs(p, v);
    else *p = v;
}

// sweep 1: a cell's neighbourhood as the TMA windows hold it. Window COMP holds, pull_x(COMP) beside the cell's own
// place, the value the cell pulls for COMP; anything else a wall cell asks for (its own populations, EAST_NOSLIP's (-1, 0, 1) entry)
// comes from the grid.
template<int WS>
struct WindowHood {
    const float *cell;  // the cell's place in window 0 of this plane's stage
    GridHood grid;
    template<int X, int Y, int Z, int COMP>
    __device__ __forceinline__ float get() const
    {
        if constexpr (X == pull_x(COMP) && Y == pull_y(COMP) && Z == pull_z(COMP)) return cell[COMP * WS + X];
        else return grid.template get<X, Y, Z, COMP>();
    }
};

// sweep 2: the intermediate level around one cell of plane q: planes q - 1 (travelling-up populations only), q, q + 1
template<int PS>
struct TileHood {
    const float *cell;  // the cell's place in plane slot 0
    int up[3];          // element offsets of planes q - 1, q, q + 1 inside a ring of three
    int flat[2];        // element offsets of planes q, q + 1 inside a ring of two
    template<int X, int Y, int Z, int COMP>
    __device__ __forceinline__ float get() const
    {
        static_assert(Z >= 0 || travels_up(COMP), "plane q - 1 only keeps the populations that travel up");
        return cell[ring_base(COMP) * PS + (travels_up(COMP) ? up[Z + 1] : flat[Z]) + X + Y * ROW];
    }
};

template<int NR>
struct Shape {
    static constexpr int PS = NR * ROW;                               // elements per plane slot of the intermediate level
    static constexpr int WS = (NR * WROW * 4 + 127) / 128 * 128 / 4;  // elements per TMA window (128-byte aligned)
    static constexpr int STAGE = WINDOWS * WS;                        // elements per stage
    static constexpr uint32_t STAGE_TX = WINDOWS * NR * WROW * 4;     // bytes the TMA unit delivers per stage
    static constexpr size_t smem(int nst) { return (size_t)nst * STAGE * 4 + (size_t)SLOTS * PS * 4 + nst * 8; }
};

// NR rows of the intermediate level per CTA, NST stages of TMA windows, MINB CTAs per SM
template<bool MACRO, int NR, int NST, int MINB>
__global__ void __launch_bounds__((NR + 1) * 32, MINB)
lbm_tb2_kernel(const __grid_constant__ CUtensorMap tmap, const float *__restrict__ src, float *__restrict__ dst, int64_t pitch,
               int64_t plane, int64_t mstride, Box box, int xa, Limits lim, const __grid_constant__ EdgeCell edge, int zchunk, int pad_x,
               int pad_y, int pad_z, int hints)
{
    static_assert(NR >= 3 && NR <= 16, "the ring cells of NR rows are one warp's work");
    typedef Shape<NR> SH;
    constexpr int PS = SH::PS, WS = SH::WS;
    extern __shared__ __align__(128) float smem[];
    float *stage = smem;                         // [NST][WINDOWS][WS]
    float *tile = smem + NST * SH::STAGE;        // [SLOTS][NR][ROW]
    uint64_t *bars = reinterpret_cast<uint64_t *>(tile + SLOTS * PS);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int X0 = xa + blockIdx.x * 32, Y0 = box.y0 + blockIdx.y * (NR - 2);
    const int zb = box.z0 + blockIdx.z * zchunk, ze = min(zb + zchunk, box.z1);
    const int first = zb - 1, planes = ze - zb + 2;  // sweep 1 covers planes zb - 1 .. ze

    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; ++s) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // hints (tuning, "lbm.tb_hints"): 1 = the results are stored with the streaming hint (first out of L2), 2 = the windows
    // are loaded with the evict-last policy (the sectors two CTAs share stay until the second one has come by)
    const bool stream_out = hints & 1;
    uint64_t keep = 0;
    if (hints & 2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(keep));
    // the windows of sweep-1 plane `first + j` go to stage j % NST
    auto issue = [&](int j) {
        float *st = stage + (j % NST) * SH::STAGE;
        uint64_t *bar = &bars[j % NST];
        mbar_expect_tx(bar, SH::STAGE_TX);
        const int cx = X0 - WLEAD + pad_x, cy = Y0 - 1 + pad_y, cz = first + j + pad_z;
        if (hints & 2) {
#pragma unroll
            for (int m = 0; m < 19; ++m) tma_load_window_hint(st + m * WS, &tmap, bar, cx, cy + pull_y(m), cz + pull_z(m), m, keep);
            tma_load_window_hint(st + 19 * WS, &tmap, bar, cx, cy, cz, STATE, keep);
        } else {
#pragma unroll
            for (int m = 0; m < 19; ++m) tma_load_window(st + m * WS, &tmap, bar, cx, cy + pull_y(m), cz + pull_z(m), m);
            tma_load_window(st + 19 * WS, &tmap, bar, cx, cy, cz, STATE);
        }
    };
    // the producer is the first thread of the ring warp: that warp has nothing to do in sweep 2, so refilling a stage
    // (20 TMA instructions) is not on the path of a warp the others wait for at the barrier
    const bool producer = threadIdx.x == NR * 32;
    if (producer)
        for (int j = 0; j < NST && j < planes; ++j) issue(j);

    // this thread's cell of the intermediate level: (row, col) in the tile, (x, y) in the grid
    const bool ring = warp == NR;
    const int row = ring ? lane >> 1 : warp;
    const int col = ring ? ((lane & 1) ? ROW - 1 : 0) : lane + 1;
    const int x = X0 - 1 + col, y = Y0 - 1 + row;
    const bool act1 = row < NR && x >= box.x0 - 1 && x <= box.x1 && y >= box.y0 - 1 && y <= box.y1;
    const bool outside_xy = x < lim.lo[0] || x >= lim.hi[0] || y < lim.lo[1] || y >= lim.hi[1];
    // ... and of the second sweep (same column: its state one plane down is the one sweep 1 saw an iteration ago)
    const bool act2 = !ring && row >= 1 && row <= NR - 2 && x >= box.x0 && x < box.x1 && y < box.y1;

    float *const mine = tile + row * ROW + col;
    const float *const window = stage + (row < NR ? row : 0) * WROW + (WLEAD - 1) + col;
    const int64_t ixy = (int64_t)y * pitch + x;

    int state = LIQUID, state_below = LIQUID;  // of this thread's cell in planes p and p - 1
    int k3 = 0;                                // k % 3
    for (int k = 0; k < planes; ++k) {
        const int p = first + k;
        const int r3 = k3 * PS, r2 = (k & 1) * PS;
        // ---- sweep 1, plane p, out of the TMA windows
        {
            uint64_t *bar = &bars[k % NST];
            const uint32_t parity = (k / NST) & 1;
            while (!mbar_try_wait(bar, parity)) {}
        }
        if (act1) {
            float out[19];
            if (outside_xy || p < lim.lo[2] || p >= lim.hi[2]) {
#pragma unroll
                for (int m = 0; m < 19; ++m) out[m] = edge.f[m];
            } else {
                const WindowHood<WS> hood = {window + (k % NST) * SH::STAGE, {src, (int64_t)p * plane + ixy, pitch, plane, mstride}};
                state = __float_as_int(hood.cell[19 * WS]);
                if (state != LIQUID) {
                    wall(state, hood, out);
                } else {
                    Pulled in;
                    pull(in, hood);
                    float rho, velX, velY, velZ;
                    liquid(in, out, rho, velX, velY, velZ);
                }
            }
#pragma unroll
            for (int m = 0; m < 19; ++m) mine[ring_base(m) * PS + (travels_up(m) ? r3 : r2)] = out[m];
        }
        __syncthreads();
        // everybody is done with this stage: refill it with the windows of plane p + NST
        if (producer && k + NST < planes) issue(k + NST);
        // ---- sweep 2, plane p - 1, out of the tile
        if (act2 && p > zb) {
            TileHood<PS> hood;
            hood.cell = mine;
            hood.up[2] = r3;                                // plane p
            hood.up[1] = (k3 == 0 ? 2 : k3 - 1) * PS;       // plane p - 1
            hood.up[0] = (k3 == 2 ? 0 : k3 + 1) * PS;       // plane p - 2
            hood.flat[1] = r2;
            hood.flat[0] = PS - r2;
            float out[19];
            const int64_t i = (int64_t)(p - 1) * plane + ixy;
            if (state_below != LIQUID) {
                wall(state_below, hood, out);
            } else {
                Pulled in;
                pull(in, hood);
                float rho, velX, velY, velZ;
                liquid(in, out, rho, velX, velY, velZ);
                if (MACRO) {
                    store_out(dst + (int64_t)DENSITY * mstride + i, rho, stream_out);
                    store_out(dst + (int64_t)VELX * mstride + i, velX, stream_out);
                    store_out(dst + (int64_t)VELY * mstride + i, velY, stream_out);
                    store_out(dst + (int64_t)VELZ * mstride + i, velZ, stream_out);
                }
            }
#pragma unroll
            for (int m = 0; m < 19; ++m) store_out(dst + (int64_t)m * mstride + i, out[m], stream_out);
        }
        state_below = state;
        k3 = k3 == 2 ? 0 : k3 + 1;
        __syncthreads();
    }
}

// ---- the same two sweeps with the two levels on DIFFERENT WARPS (the default): no CTA-wide barrier per plane.
//
// In lbm_tb2_kernel all warps compute sweep 1 of a plane, meet at a barrier, compute sweep 2, meet again: the issue
// slots are 45 % busy, 28 % of the warp time is spent at the barriers (profiles/r4a_r4h). Here NR + 1 warps do
// nothing but sweep 1 (plane after plane, out of the TMA windows into the rings), NR - 2 warps nothing but sweep 2
// (out of the rings into the grid), one warp refills the TMA stages; they meet only through mbarriers:
//   tma_full[s]  : the windows of a plane have landed                 (TMA unit -> sweep-1 warps)
//   tma_empty[s] : every sweep-1 warp has read its row of the windows  (sweep-1 warps -> producer)
//   l1_full[k%4] : sweep-1 plane k is in the rings                     (sweep-1 warps -> sweep-2 warps)
//   l2_done[j%4] : sweep-2 plane j has been read out of the rings      (sweep-2 warps -> sweep-1 warps)
// The rings are one plane deeper than in the kernel above (4 for the populations that travel up, 3 for the others, 3 for
// the state), so sweep 1 may run two to three planes ahead of sweep 2 and neither waits for the other's slowest warp.
constexpr int WSLOTS = 5 * 4 + 14 * 3 + 3;

__host__ __device__ constexpr int wring_base(int comp)
{
    return comp == T ? 0 : comp == TW ? 4 : comp == TE ? 8 : comp == TN ? 12 : comp == TS ? 16 :
           20 + 3 * (comp == C ? 0 : comp == N ? 1 : comp == E ? 2 : comp == W ? 3 : comp == S ? 4 : comp == B ? 5 :
                     comp == NW ? 6 : comp == SW ? 7 : comp == NE ? 8 : comp == SE ? 9 : comp == BW ? 10 : comp == BE ? 11 :
                     comp == BN ? 12 : comp == BS ? 13 : 14);   // 14: the state
}

template<int PS>
struct WideTileHood {
    const float *cell;  // the cell's place in plane slot 0
    int up[3];          // element offsets of planes q - 1, q, q + 1 inside a ring of four
    int flat[2];        // element offsets of planes q, q + 1 inside a ring of three
    template<int X, int Y, int Z, int COMP>
    __device__ __forceinline__ float get() const
    {
        static_assert(Z >= 0 || travels_up(COMP), "plane q - 1 only keeps the populations that travel up");
        return cell[wring_base(COMP) * PS + (travels_up(COMP) ? up[Z + 1] : flat[Z]) + X + Y * ROW];
    }
};

__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {}
}

template<int NR>
struct WideShape {
    static constexpr int PS = NR * ROW;
    static constexpr int WS = Shape<NR>::WS;
    static constexpr int STAGE = Shape<NR>::STAGE;
    static constexpr int WARPS = 2 * NR;   // NR + 1 for sweep 1, NR - 2 for sweep 2, one producer
    static constexpr size_t smem(int nst) { return (size_t)nst * STAGE * 4 + (size_t)WSLOTS * PS * 4 + (2 * nst + 8) * 8 + 8; }
};

template<bool MACRO, int NR, int NST>
__global__ void __launch_bounds__(2 * NR * 32, 1)
lbm_tb2w_kernel(const __grid_constant__ CUtensorMap tmap, const float *__restrict__ src, float *__restrict__ dst, int64_t pitch,
                int64_t plane, int64_t mstride, Box box, int xa, Limits lim, const __grid_constant__ EdgeCell edge, int zchunk, int pad_x,
                int pad_y, int pad_z)
{
    static_assert(NR >= 3 && NR <= 16, "the ring cells of NR rows are one warp's work");
    typedef WideShape<NR> SH;
    constexpr int PS = SH::PS, WS = SH::WS;
    extern __shared__ __align__(128) float smem[];
    float *stage = smem;                         // [NST][WINDOWS][WS]
    float *tile = smem + NST * SH::STAGE;        // [WSLOTS][NR][ROW]
    uint64_t *tma_full = reinterpret_cast<uint64_t *>(tile + WSLOTS * PS);
    uint64_t *tma_empty = tma_full + NST;
    uint64_t *l1_full = tma_empty + NST;         // [4]
    uint64_t *l2_done = l1_full + 4;             // [4]
    int *sink = reinterpret_cast<int *>(l2_done + 4);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int X0 = xa + blockIdx.x * 32, Y0 = box.y0 + blockIdx.y * (NR - 2);
    const int zb = box.z0 + blockIdx.z * zchunk, ze = min(zb + zchunk, box.z1);
    const int first = zb - 1, planes = ze - zb + 2;  // sweep 1 covers planes first + k, k = 0 .. planes - 1

    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; ++s) {
            mbar_init(&tma_full[s], 1);
            mbar_init(&tma_empty[s], NR + 1);
        }
        for (int i = 0; i < 4; ++i) {
            mbar_init(&l1_full[i], NR + 1);
            mbar_init(&l2_done[i], NR - 2);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == 2 * NR - 1) {
        // ---- the producer: refills stage k % NST with the windows of plane k as soon as the sweep-1 warps have read it
        if (lane == 0) {
            for (int k = 0; k < planes; ++k) {
                const int s = k % NST;
                if (k >= NST) mbar_wait(&tma_empty[s], ((k / NST) - 1) & 1);
                float *st = stage + s * SH::STAGE;
                mbar_expect_tx(&tma_full[s], Shape<NR>::STAGE_TX);
                const int cx = X0 - WLEAD + pad_x, cy = Y0 - 1 + pad_y, cz = first + k + pad_z;
#pragma unroll
                for (int m = 0; m < 19; ++m) tma_load_window(st + m * WS, &tmap, &tma_full[s], cx, cy + pull_y(m), cz + pull_z(m), m);
                tma_load_window(st + 19 * WS, &tmap, &tma_full[s], cx, cy, cz, STATE);
            }
        }
        return;
    }

    if (warp <= NR) {
        // ---- sweep 1: plane after plane out of the TMA windows into the rings
        const bool ring = warp == NR;
        const int row = ring ? lane >> 1 : warp;
        const int col = ring ? ((lane & 1) ? ROW - 1 : 0) : lane + 1;
        const int x = X0 - 1 + col, y = Y0 - 1 + row;
        const bool act1 = row < NR && x >= box.x0 - 1 && x <= box.x1 && y >= box.y0 - 1 && y <= box.y1;
        const bool outside_xy = x < lim.lo[0] || x >= lim.hi[0] || y < lim.lo[1] || y >= lim.hi[1];
        float *const mine = tile + row * ROW + col;
        const float *const window = stage + (row < NR ? row : 0) * WROW + (WLEAD - 1) + col;
        int k3 = 0, k4 = 0, ks = 0;   // k % 3, k % 4, k % NST
        for (int k = 0; k < planes; ++k) {
            const int p = first + k;
            // the ring slots of plane k held plane k - 4 / k - 3: sweep 2 must be through with plane k - 3
            if (k >= 3) mbar_wait(&l2_done[(k4 + 1) & 3], ((k - 3) >> 2) & 1);
            mbar_wait(&tma_full[ks], (k / NST) & 1);
            float out[19];
            int state = LIQUID;
            int seen = 0, all_loaded = 0;
            if (act1) {
                if (outside_xy || p < lim.lo[2] || p >= lim.hi[2]) {
#pragma unroll
                    for (int m = 0; m < 19; ++m) out[m] = edge.f[m];
                } else {
                    const float *cell = window + ks * SH::STAGE;
                    state = __float_as_int(cell[19 * WS]);
                    if (state != LIQUID) {
                        // a wall cell reads its own populations from the grid (faces only). The strides go through an
                        // empty asm so that the 19 member addresses are worked out HERE, not hoisted out of the plane loop
                        // into registers every liquid cell would pay for
                        int64_t ms = mstride, pl = plane, pi = pitch;
                        asm volatile("" : "+l"(ms), "+l"(pl), "+l"(pi));
                        const WindowHood<WS> hood = {cell, {src, (int64_t)p * pl + (int64_t)y * pi + x, pi, pl, ms}};
                        wall(state, hood, out);
                        // The populations a wall cell bounces back go from the windows straight into `out`: no
                        // arithmetic needs them. Fold them into `seen` (see the arrival on tma_empty below).
#pragma unroll
                        for (int m = 0; m < 19; ++m) seen ^= __float_as_int(out[m]);
                    } else {
                        const WindowHood<WS> hood = {cell, {src, 0, 0, 0, 0}};
                        Pulled in;
                        pull(in, hood);
                        float rho, velX, velY, velZ;
                        liquid(in, out, rho, velX, velY, velZ);
                    }
                }
            }
            // This warp has read its row of the windows — but "has issued the loads" is not enough: the arrival lets the
            // producer refill the stage, and a shared-memory load that is still in flight then returns the plane after
            // next (one row beside a wall plane was wrong in about one launch of twenty at 512^3 that way, profiles/
            // r4o_r4u: a wall cell's bounced-back populations feed no arithmetic). So an instruction that NEEDS every
            // loaded value is put in front of the arrival — warps issue in order, and a warp-wide instruction waits for
            // the producers of its operands in all lanes: `all_loaded` is known when out[C] (liquid cells: C's new value
            // needs the density, i.e. all 19 pulled values) and `seen` (wall cells: everything they loaded) are. The
            // store it guards goes to a word nobody reads and practically never happens.
            if (act1) all_loaded = seen ^ __float_as_int(out[C]);
            if (all_loaded == 0x7fc5a5a5) *sink = 1;
            __syncwarp();
            if (lane == 0) mbar_arrive(&tma_empty[ks]);
            if (act1) {
                const int r4 = k4 * PS, r3 = k3 * PS;
#pragma unroll
                for (int m = 0; m < 19; ++m) mine[wring_base(m) * PS + (travels_up(m) ? r4 : r3)] = out[m];
                mine[wring_base(19) * PS + r3] = __int_as_float(state);
            }
            // the ring stores of all lanes are visible before lane 0 announces the plane
            __threadfence_block();
            __syncwarp();
            if (lane == 0) mbar_arrive(&l1_full[k4]);
            k3 = k3 == 2 ? 0 : k3 + 1;
            k4 = (k4 + 1) & 3;
            ks = ks == NST - 1 ? 0 : ks + 1;
        }
        return;
    }

    // ---- sweep 2: plane j = k (ring index) out of the rings into the grid, for k = 1 .. planes - 2
    {
        const int row = warp - NR;   // 1 .. NR - 2
        const int col = lane + 1;
        const int x = X0 - 1 + col, y = Y0 - 1 + row;
        const bool act2 = x >= box.x0 && x < box.x1 && y < box.y1;
        const float *const mine = tile + row * ROW + col;
        const int64_t ixy = (int64_t)y * pitch + x;
        // plane 0 is nobody's second sweep: its slot of l2_done is completed right away so that parities stay uniform
        if (lane == 0) mbar_arrive(&l2_done[0]);
        int j3 = 1, j4 = 1;
        for (int j = 1; j <= planes - 2; ++j) {
            mbar_wait(&l1_full[(j4 + 1) & 3], ((j + 1) >> 2) & 1);   // planes j - 1, j, j + 1 of sweep 1 are there
            float out[19];
            const int64_t i = (int64_t)(first + j) * plane + ixy;
            if (act2) {
                WideTileHood<PS> hood;
                hood.cell = mine;
                hood.up[0] = ((j4 + 3) & 3) * PS;
                hood.up[1] = j4 * PS;
                hood.up[2] = ((j4 + 1) & 3) * PS;
                hood.flat[0] = j3 * PS;
                hood.flat[1] = (j3 == 2 ? 0 : j3 + 1) * PS;
                const int state = __float_as_int(mine[wring_base(19) * PS + j3 * PS]);
                if (state != LIQUID) {
                    wall(state, hood, out);
                } else {
                    Pulled in;
                    pull(in, hood);
                    float rho, velX, velY, velZ;
                    liquid(in, out, rho, velX, velY, velZ);
                    if (MACRO) {
                        dst[(int64_t)DENSITY * mstride + i] = rho;
                        dst[(int64_t)VELX * mstride + i] = velX;
                        dst[(int64_t)VELY * mstride + i] = velY;
                        dst[(int64_t)VELZ * mstride + i] = velZ;
                    }
                }
            }
            if (act2) {
#pragma unroll
                for (int m = 0; m < 19; ++m) dst[(int64_t)m * mstride + i] = out[m];
            }
            // the arrival comes after the stores: they need every value this plane took from the rings (a wall cell copies
            // its own populations from the rings straight into `out`), so all those loads have been performed
            __syncwarp();
            if (lane == 0) mbar_arrive(&l2_done[j4]);
            j3 = j3 == 2 ? 0 : j3 + 1;
            j4 = (j4 + 1) & 3;
        }
    }
}

typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// TMA descriptor of the current buffer: the 24 padded member arrays as ONE rank-4 tensor of 4-byte elements
// (x, y, z, member; INT32: the windows are copied bit for bit, the state is an integer), box = one 40 x NR window of one member's plane. Coordinates outside the arrays are zero-filled
// (only cells nobody uses pull from there).
int window_map(b200geo_grid *g, int rows, CUtensorMap *out)
{
    EncodeTiled enc = (EncodeTiled)tensor_map_encoder();
    if (!enc) return fail(B200GEO_ERR_CUDA, "CUDA error: cuTensorMapEncodeTiled is not available in this driver");
    const MemberLayout& L = g->m[0];
    cuuint64_t dims[4] = {(cuuint64_t)L.pitch, (cuuint64_t)(g->d[1] + 2 * g->g[1]), (cuuint64_t)(g->d[2] + 2 * g->g[2]), 24};
    cuuint64_t strides[3] = {(cuuint64_t)L.pitch * 4, (cuuint64_t)L.plane * 4, (cuuint64_t)g->m[1].offset};
    cuuint32_t boxdim[4] = {(cuuint32_t)WROW, (cuuint32_t)rows, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUtensorMapL2promotion promo = g_tuning.lbm_tb_promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE :
        g_tuning.lbm_tb_promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B :
        g_tuning.lbm_tb_promo == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    const char *dbg = getenv("B200GEO_LBM_TMA_F32");
    CUresult r = enc(out, dbg ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_INT32, 4, g->member_ptr(0, 0), dims, strides, boxdim, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(B200GEO_ERR_CUDA, "CUDA error: cuTensorMapEncodeTiled failed");
    return B200GEO_OK;
}

template<int NR, int NST, int MINB>
int launch_tb2(b200geo_grid *g, const Box& box, const Limits& lim, bool store_macroscopic, cudaStream_t s)
{
    const MemberLayout& L = g->m[0];
    for (int m = 1; m < 24; ++m)
        if (g->m[m].offset != (int64_t)m * g->m[1].offset) return fail(B200GEO_ERR_LOGIC, "LBM member arrays are not equally spaced");
    const int64_t mstride = g->m[1].offset / 4;
    const float *src = (const float *)g->member_ptr(0, 0) + L.origin;
    float *dst = (float *)g->member_ptr(0, 1) + L.origin;
    const int ny = box.y1 - box.y0, nz = box.z1 - box.z0;
    const int xa = box.x0 & ~3;  // tiles start on multiples of 16 bytes (TMA), whole 128-byte lines for aligned boxes
    const int gx = (box.x1 - xa + 31) / 32, gy = (ny + NR - 3) / (NR - 2);
    // long z chunks amortise the extra sweep-1 planes at either end (2 of them: 3 % at 64 planes); enough chunks to fill
    // the machine several times over
    int zchunk = g_tuning.lbm_tb_zchunk > 0 ? g_tuning.lbm_tb_zchunk : 64;
    while (zchunk > 8 && (int64_t)gx * gy * ((nz + zchunk - 1) / zchunk) < 148 * MINB * 4) zchunk /= 2;
    const int gz = (nz + zchunk - 1) / zchunk;
    if (gy > 65535 || gz > 65535) return fail(B200GEO_ERR_OUT_OF_RANGE, "grid dimension too large");
    CUtensorMap map;
    int rc = window_map(g, NR, &map);
    if (rc) return rc;
    EdgeCell edge;
    for (int m = 0; m < 19; ++m) memcpy(&edge.f[m], g->edge + g->m[m].edge_offset, 4);
    const size_t smem = Shape<NR>::smem(NST);
    static bool attr_set[2][64] = {{false}};
    auto k1 = lbm_tb2_kernel<true, NR, NST, MINB>;
    auto k0 = lbm_tb2_kernel<false, NR, NST, MINB>;
    const int which = store_macroscopic ? 1 : 0;
    if (g->device < 0 || g->device >= 64 || !attr_set[which][g->device]) {
        B200GEO_CUDA(cudaFuncSetAttribute(which ? k1 : k0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (g->device >= 0 && g->device < 64) attr_set[which][g->device] = true;
    }
    dim3 grid(gx, gy, gz);
    (which ? k1 : k0)<<<grid, (NR + 1) * 32, smem, s>>>(map, src, dst, L.pitch, L.plane, mstride, box, xa, lim, edge, zchunk, L.lead,
                                                       g->g[1], g->g[2], g_tuning.lbm_tb_hints);
    count_launch();
    return check_cuda(cudaGetLastError(), "fused lbm sweeps");
}

template<int NR, int NST>
int launch_tb2w(b200geo_grid *g, const Box& box, const Limits& lim, bool store_macroscopic, cudaStream_t s)
{
    const MemberLayout& L = g->m[0];
    for (int m = 1; m < 24; ++m)
        if (g->m[m].offset != (int64_t)m * g->m[1].offset) return fail(B200GEO_ERR_LOGIC, "LBM member arrays are not equally spaced");
    const int64_t mstride = g->m[1].offset / 4;
    const float *src = (const float *)g->member_ptr(0, 0) + L.origin;
    float *dst = (float *)g->member_ptr(0, 1) + L.origin;
    const int ny = box.y1 - box.y0, nz = box.z1 - box.z0;
    const int xa = box.x0 & ~3;
    const int gx = (box.x1 - xa + 31) / 32, gy = (ny + NR - 3) / (NR - 2);
    // short z chunks: neighbouring columns start together more often, so more of the rim rows they share are still in L2
    // (13.7 GB read per launch at 512^3 with 32 planes, 15.2 GB with 64; profiles/r4f) — worth the two extra sweep-1 planes
    int zchunk = g_tuning.lbm_tb_zchunk > 0 ? g_tuning.lbm_tb_zchunk : 32;
    while (zchunk > 8 && (int64_t)gx * gy * ((nz + zchunk - 1) / zchunk) < 148 * 4) zchunk /= 2;
    const int gz = (nz + zchunk - 1) / zchunk;
    if (gy > 65535 || gz > 65535) return fail(B200GEO_ERR_OUT_OF_RANGE, "grid dimension too large");
    CUtensorMap map;
    int rc = window_map(g, NR, &map);
    if (rc) return rc;
    EdgeCell edge;
    for (int m = 0; m < 19; ++m) memcpy(&edge.f[m], g->edge + g->m[m].edge_offset, 4);
    const size_t smem = WideShape<NR>::smem(NST);
    static bool attr_set[2][64] = {{false}};
    auto k1 = lbm_tb2w_kernel<true, NR, NST>;
    auto k0 = lbm_tb2w_kernel<false, NR, NST>;
    const int which = store_macroscopic ? 1 : 0;
    if (g->device < 0 || g->device >= 64 || !attr_set[which][g->device]) {
        B200GEO_CUDA(cudaFuncSetAttribute(which ? k1 : k0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (g->device >= 0 && g->device < 64) attr_set[which][g->device] = true;
    }
    dim3 grid(gx, gy, gz);
    (which ? k1 : k0)<<<grid, 2 * NR * 32, smem, s>>>(map, src, dst, L.pitch, L.plane, mstride, box, xa, lim, edge, zchunk, L.lead,
                                                      g->g[1], g->g[2]);
    count_launch();
    return check_cuda(cudaGetLastError(), "fused lbm sweeps");
}

}

// Two sweeps over `box`; the grid's current buffer must hold valid cells two deep around it wherever the box does
// not end at a Cube boundary (the caller checks: b200geo_step, b200geo_update_box_n).
int sweep_lbm_tb2(b200geo_grid *g, const Box& box, bool store_macroscopic, cudaStream_t s)
{
    Limits lim;
    for (int i = 0; i < 3; ++i) {
        lim.lo[i] = g->desc.ghost_mode[i][0] == B200GEO_GHOST_EDGE ? 0 : INT_MIN;
        lim.hi[i] = g->desc.ghost_mode[i][1] == B200GEO_GHOST_EDGE ? g->d[i] : INT_MAX;
    }
    if (g_tuning.lbm_tb_warps) {   // sweep 1 and sweep 2 on different warps
        // rows of the intermediate level x TMA stages: what fits into 227 KB beside rings of 65 plane slots. Three stages
        // beat two wider tiles (59 against 54 GLUPS at 512^3, profiles/r4t): the launch waits for its windows
        switch (g_tuning.lbm_tb_rows) {
        case 142: return launch_tb2w<14, 2>(g, box, lim, store_macroscopic, s);
        case 122: return launch_tb2w<12, 2>(g, box, lim, store_macroscopic, s);
        case 113: return launch_tb2w<11, 3>(g, box, lim, store_macroscopic, s);
        case 104: return launch_tb2w<10, 4>(g, box, lim, store_macroscopic, s);
        case 103: return launch_tb2w<10, 3>(g, box, lim, store_macroscopic, s);
        default: return launch_tb2w<12, 3>(g, box, lim, store_macroscopic, s);
        }
    }
    switch (g_tuning.lbm_tb_rows) {
    case 8: return launch_tb2<8, 2, 2>(g, box, lim, store_macroscopic, s);     // two small CTAs per SM
    case 16: return launch_tb2<16, 2, 1>(g, box, lim, store_macroscopic, s);   // 186 KB
    case 142: return launch_tb2<14, 2, 1>(g, box, lim, store_macroscopic, s);
    default: return launch_tb2<14, 3, 1>(g, box, lim, store_macroscopic, s);   // 14 rows, three stages: 203 KB
    }
}

}
