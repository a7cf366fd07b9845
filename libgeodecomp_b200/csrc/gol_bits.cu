// Bit-packed Game of Life: many sweeps per HBM round trip for the 1-byte-cell grid.
//
// The interface keeps the reference's cell (ConwayCell: one `bool alive`, src/examples/gameoflife/
// main.cpp:25-62): the grid in HBM, its ghost ring, region and member I/O and the halo exchange all
// stay byte based. When b200geo_step is asked for several sweeps at once, the current buffer is
// packed to one bit per cell, the sweeps run on the packed copy (16384^2 cells = 32 MiB per buffer:
// both buffers live in the 126 MB L2, so between pack and unpack no sweep touches HBM at all), and
// the result is unpacked into the byte grid. Same rule, integer logic only: bit-exact.
//
// Sweep kernel: a lane owns one 32-cell word column and marches down the rows. West/east neighbours
// are funnel shifts of the word with its lane neighbours' words (warp shuffles; the two border
// lanes of a warp read the adjacent word themselves). Per row the three-cell sums of the row
// (above/below use) and the two-cell sum (centre use) are kept as bit planes; a new word is
//   n = up3 + mid2 + down3 (bit-sliced adders), alive' = (n == 3) | (alive & n == 2).
// About 25 logic instructions per 32 cells and sweep.
//
// Boundaries need no ghost storage: rows / words outside the grid read as the constant edge word
// (Cube) or wrap (Torus, widths that are a multiple of 32 cells); the unused high bits of a ragged
// last word are kept equal to the edge value.
#include "grid.h"

#include <cstring>

namespace b200geo {

namespace {

struct BitGrid {
    int W, ny;            // words per row, rows
    int wrap_x, wrap_y;
    uint32_t edge;        // 0 or ~0: the constant edge cell of a Cube axis
    uint32_t last_mask;   // valid bits of the last word of a row
};

__device__ __forceinline__ uint32_t xor3(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}

__device__ __forceinline__ uint32_t maj3(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xe8;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}

struct RowSums {
    uint32_t c;        // the row's own word
    uint32_t t0, t1;   // west + centre + east, bit planes
    uint32_t p0, p1;   // west + east
};

struct RawRow {
    uint32_t c, l, r;   // own word; west / east words (meaningful in a warp's border lanes only)
};

// what a lane needs to know about its column, computed once
struct Column {
    int w;
    bool active;            // w < W
    bool border_l, border_r;  // the lane takes its west / east word from memory instead of a shuffle
    int off_l, off_r;       // word offsets of those loads relative to the lane's own word; 0 = edge word
};

// issue the loads of one row (no use of the values yet: several rows are kept in flight).
// SAFE_Y: the caller guarantees 0 <= y < ny, and `row` already points at the lane's word of that row.
template<bool SAFE_Y>
__device__ __forceinline__ RawRow fetch_row(const uint32_t *__restrict__ src, const uint32_t *row, const BitGrid& B,
                                            const Column& col, int y)
{
    RawRow q;
    q.c = q.l = q.r = B.edge;
    if (!SAFE_Y) {
        if (y < 0 || y >= B.ny) {
            if (!B.wrap_y) return q;
            y = y < 0 ? y + B.ny : y - B.ny;
        }
        row = src + (int64_t)y * B.W + col.w;
    }
    if (col.active) q.c = *row;
    if (col.border_l && col.off_l) q.l = row[col.off_l];
    if (col.border_r && col.off_r) q.r = row[col.off_r];
    return q;
}

// all 32 lanes of the warp call this together
__device__ __forceinline__ RowSums combine_row(const RawRow& q, const Column& col)
{
    uint32_t left = __shfl_up_sync(0xffffffffu, q.c, 1);
    uint32_t right = __shfl_down_sync(0xffffffffu, q.c, 1);
    if (col.border_l) left = q.l;
    if (col.border_r) right = q.r;
    const uint32_t l = __funnelshift_l(left, q.c, 1);    // bit i <- cell i - 1
    const uint32_t r = __funnelshift_r(q.c, right, 1);   // bit i <- cell i + 1
    RowSums s;
    s.c = q.c;
    s.t0 = xor3(l, q.c, r);
    s.t1 = maj3(l, q.c, r);
    s.p0 = l ^ r;
    s.p1 = l & r;
    return s;
}

constexpr int AHEAD = 4;   // rows in flight per lane

template<bool SAFE_Y>
__device__ __forceinline__ void march(const uint32_t *__restrict__ src, uint32_t *__restrict__ dst, const BitGrid& B,
                                      const Column& col, int yb, int ye)
{
    const uint32_t *row = src + (int64_t)(yb - 1) * B.W + col.w;   // only dereferenced when SAFE_Y
    uint32_t *out = dst + (int64_t)yb * B.W + col.w;
    RawRow q[AHEAD];
    RowSums up = combine_row(fetch_row<SAFE_Y>(src, row, B, col, yb - 1), col);
    RowSums mid = combine_row(fetch_row<SAFE_Y>(src, row + B.W, B, col, yb), col);
    row += 2 * B.W;
#pragma unroll
    for (int k = 0; k < AHEAD; ++k, row += B.W) q[k] = fetch_row<SAFE_Y>(src, row, B, col, yb + 1 + k);
    const bool last_word = col.w == B.W - 1;
    for (int y = yb; y < ye; y += AHEAD) {
#pragma unroll
        for (int k = 0; k < AHEAD; ++k) {
            if (y + k >= ye) break;
            const RowSums dn = combine_row(q[k], col);
            q[k] = fetch_row<SAFE_Y>(src, row, B, col, y + k + 1 + AHEAD);
            row += B.W;
            // n = up.t + dn.t + mid.p, bit sliced: 0..8 in planes n0..n3
            const uint32_t n0 = xor3(up.t0, dn.t0, mid.p0);
            const uint32_t k0 = maj3(up.t0, dn.t0, mid.p0);
            const uint32_t a = xor3(up.t1, dn.t1, mid.p1);
            const uint32_t k1 = maj3(up.t1, dn.t1, mid.p1);
            const uint32_t n1 = a ^ k0;
            const uint32_t k2 = a & k0;
            const uint32_t n2 = k1 ^ k2;
            const uint32_t n3 = k1 & k2;
            // alive' = (n == 3) | (alive & n == 2) = n1 & ~n2 & ~n3 & (n0 | alive)
            uint32_t next = n1 & ~(n2 | n3) & (n0 | mid.c);
            if (last_word) next = (next & B.last_mask) | (B.edge & ~B.last_mask);
            if (col.active) *out = next;
            out += B.W;
            up = mid;
            mid = dn;
        }
    }
}

__global__ void __launch_bounds__(128)
gol_bits_kernel(const uint32_t *__restrict__ src, uint32_t *__restrict__ dst, BitGrid B, int rows_per_block)
{
    // programmatic dependent launch: a packed sweep (~20 us) is short next to the gap between two dependent
    // launches, so sweep n + 1 is scheduled while sweep n drains; no memory access before the wait
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int lane = threadIdx.x & 31;
    Column col;
    col.w = blockIdx.x * blockDim.x + threadIdx.x;
    const int yb = blockIdx.y * rows_per_block;
    const int ye = min(yb + rows_per_block, B.ny);
    // whole warps only: the shuffles need every lane (blockDim.x is a multiple of 32 and a warp
    // whose first word is outside the grid has nothing to do)
    if (col.w - lane >= B.W) return;
    col.active = col.w < B.W;
    col.border_l = lane == 0 || col.w == 0;
    col.border_r = lane == 31 || col.w >= B.W - 1;
    col.off_l = col.w > 0 ? -1 : (B.wrap_x ? B.W - 1 : 0);
    col.off_r = col.w < B.W - 1 ? 1 : (B.wrap_x && col.w == B.W - 1 ? -(B.W - 1) : 0);
    if (!col.active) col.off_l = col.off_r = 0;
    // every row this CTA touches (one above, AHEAD + 1 below its last) inside the grid: pointer
    // arithmetic only; otherwise the general path with edge / wrap handling per row
    if (yb >= 1 && ye + AHEAD + 1 <= B.ny) march<true>(src, dst, B, col, yb, ye);
    else march<false>(src, dst, B, col, yb, ye);
}

// byte grid (interior, padded pitch) -> bits; a thread makes one word from 32 bytes
__global__ void gol_pack_kernel(const uint8_t *__restrict__ src, int64_t pitch, int nx, uint32_t *__restrict__ dst, BitGrid B)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (w >= B.W) return;
    const uint8_t *p = src + (int64_t)y * pitch + 32 * w;
    uint32_t bits = 0;
    if (32 * w + 32 <= nx) {
        const uint4 a = *reinterpret_cast<const uint4 *>(p), b = *reinterpret_cast<const uint4 *>(p + 16);
        const uint32_t v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) bits |= (((v[i] & 0x01010101u) * 0x01020408u) >> 24 & 0xfu) << (4 * i);
    } else {
        for (int i = 0; i < 32; ++i) {
            uint32_t cell = 32 * w + i < nx ? (p[i] & 1u) : (B.edge & 1u);
            bits |= cell << i;
        }
    }
    dst[(int64_t)y * B.W + w] = bits;
}

__global__ void gol_unpack_kernel(const uint32_t *__restrict__ src, BitGrid B, uint8_t *__restrict__ dst, int64_t pitch, int nx)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (w >= B.W) return;
    const uint32_t bits = src[(int64_t)y * B.W + w];
    uint8_t *q = dst + (int64_t)y * pitch + 32 * w;
    if (32 * w + 32 <= nx) {
        uint32_t v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = (((bits >> (4 * i)) & 0xfu) * 0x00204081u) & 0x01010101u;
        *reinterpret_cast<uint4 *>(q) = make_uint4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<uint4 *>(q + 16) = make_uint4(v[4], v[5], v[6], v[7]);
    } else {
        for (int i = 0; 32 * w + i < nx; ++i) q[i] = (uint8_t)((bits >> i) & 1u);
    }
}

}

// Can `sweeps` sweeps of this grid run on the packed copy? (2-D Game of Life grid, no PEER side,
// a Torus in x only when the width is a whole number of words.)
bool gol_bits_applicable(const b200geo_grid *g)
{
    if (g->n != 1 || g->m[0].elem != 1 || g->d[2] != 1) return false;
    for (int i = 0; i < 2; ++i)
        for (int side = 0; side < 2; ++side)
            if (g->desc.ghost_mode[i][side] == B200GEO_GHOST_PEER) return false;
    if (g->desc.ghost_mode[0][0] == B200GEO_GHOST_WRAP && g->d[0] % 32 != 0) return false;
    return true;
}

int sweep_gol_bits(b200geo_grid *g, uint32_t sweeps, cudaStream_t s)
{
    const MemberLayout& L = g->m[0];
    BitGrid B;
    B.W = (g->d[0] + 31) / 32;
    B.ny = g->d[1];
    B.wrap_x = g->desc.ghost_mode[0][0] == B200GEO_GHOST_WRAP;
    B.wrap_y = g->desc.ghost_mode[1][0] == B200GEO_GHOST_WRAP;
    B.edge = g->edge[0] ? 0xffffffffu : 0u;
    int tail = g->d[0] % 32;
    B.last_mask = tail ? (1u << tail) - 1 : 0xffffffffu;
    if (B.ny > 65535) return fail(B200GEO_ERR_OUT_OF_RANGE, "grid dimension too large");

    size_t words = (size_t)B.W * B.ny;
    size_t need = 2 * words * sizeof(uint32_t);
    if (g->bits_bytes < need) {
        if (g->bits) { cudaStreamSynchronize(s); cudaFree(g->bits); g->bits = 0; g->bits_bytes = 0; }
        cudaError_t e = cudaMalloc(&g->bits, need);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(B200GEO_ERR_NOMEM, std::string("cudaMalloc failed: ") + cudaGetErrorString(e));
        }
        g->bits_bytes = need;
    }
    uint32_t *buf[2] = {(uint32_t *)g->bits, (uint32_t *)g->bits + words};

    // 128-bit accesses of the pack / unpack kernels: interior x = 0 is 128 bytes into a row and the pitch a multiple of 128
    uint8_t *cur = (uint8_t *)g->member_ptr(0, 0) + L.origin;
    dim3 pgrid((B.W + 127) / 128, B.ny);
    gol_pack_kernel<<<pgrid, 128, 0, s>>>(cur, L.pitch, g->d[0], buf[0], B);
    count_launch();
    int rows = 32;
    int gx = (B.W + 127) / 128;
    while (rows > 4 && (int64_t)gx * ((B.ny + rows - 1) / rows) < 148 * 16) rows /= 2;
    if (g_tuning.gol_bits_rows > 0) rows = g_tuning.gol_bits_rows;
    dim3 sgrid(gx, (B.ny + rows - 1) / rows);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = sgrid;
    cfg.blockDim = dim3(128);
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    for (uint32_t t = 0; t < sweeps; ++t) {
        const uint32_t *src = buf[t & 1];
        uint32_t *dst = buf[(t + 1) & 1];
        B200GEO_CUDA(cudaLaunchKernelEx(&cfg, gol_bits_kernel, src, dst, B, rows));
        count_launch();
    }
    gol_unpack_kernel<<<pgrid, 128, 0, s>>>(buf[sweeps & 1], B, cur, L.pitch, g->d[0]);
    count_launch();
    return check_cuda(cudaGetLastError(), "bit-packed gol sweeps");
}

}
