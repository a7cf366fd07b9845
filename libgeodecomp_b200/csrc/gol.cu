// Conway's Game of Life sweep on 1-byte cells (B200GEO_KERNEL_GOL).
//
// Replaces the per-cell ConwayCell::update reached through VanillaUpdateFunctor + CoordMap
// (storage/vanillaupdatefunctor.h:12-36, storage/coordmap.h:34-43; rule:
// src/examples/gameoflife/main.cpp:36-59): nine bounds-checked loads per cell there, here a
// thread owns 16 x-adjacent cells (one 128-bit access per row) and marches down y keeping the
// horizontal 3-sums of the last three rows in registers. Cell bytes are 0/1, so four cells are
// summed at once with ordinary 32-bit adds (a neighbourhood sum is <= 9, no carry between bytes).
// Integer work: bit-exact. Roofline: HBM, 2 algorithmic bytes per update.
#include "grid.h"

namespace b200geo {

namespace {

struct Row4 { uint32_t w[4]; };

// per byte: left + centre + right
__device__ __forceinline__ Row4 hsum(const uint8_t *p, bool has_right)
{
    uint4 c = *reinterpret_cast<const uint4 *>(p);
    uint32_t l = p[-1], r = has_right ? p[16] : 0;
    uint32_t w[6] = {l << 24, c.x, c.y, c.z, c.w, r};
    Row4 h;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint32_t left = __funnelshift_l(w[i], w[i + 1], 8);    // byte k <- cell x-1
        uint32_t right = __funnelshift_r(w[i + 1], w[i + 2], 8); // byte k <- cell x+1
        h.w[i] = left + w[i + 1] + right;
    }
    return h;
}

// bit 7 of every byte set where the byte (< 0x80) is zero
__device__ __forceinline__ uint32_t zero_bytes(uint32_t v)
{
    return ~(v + 0x7f7f7f7fu) & 0x80808080u;
}

__global__ void __launch_bounds__(128)
gol_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, int64_t pitch, Box box,
           int xa, int rows_per_block, int x_limit)
{
    const int x = xa + 16 * (blockIdx.x * blockDim.x + threadIdx.x);
    if (x >= box.x1) return;
    const int yb = box.y0 + blockIdx.y * rows_per_block;
    const int ye = min(yb + rows_per_block, box.y1);
    const bool has_right = x + 16 < x_limit;
    const bool full = x >= box.x0 && x + 16 <= box.x1;

    const uint8_t *p = src + (int64_t)yb * pitch + x;
    uint8_t *q = dst + (int64_t)yb * pitch + x;
    Row4 hm = hsum(p - pitch, has_right), hc = hsum(p, has_right);
    uint4 cc = *reinterpret_cast<const uint4 *>(p);
#pragma unroll 2
    for (int y = yb; y < ye; ++y, p += pitch, q += pitch) {
        Row4 hp = hsum(p + pitch, has_right);
        uint4 cn = *reinterpret_cast<const uint4 *>(p + pitch);
        uint32_t self[4] = {cc.x, cc.y, cc.z, cc.w}, out[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint32_t n9 = hm.w[i] + hc.w[i] + hp.w[i];   // includes the cell itself
            // alive' = (n9 == 3) || (alive && n9 == 4)
            uint32_t eq3 = zero_bytes(n9 ^ 0x03030303u), eq4 = zero_bytes(n9 ^ 0x04040404u);
            out[i] = ((eq3 | (eq4 & (self[i] << 7))) >> 7) & 0x01010101u;
        }
        if (full) {
            *reinterpret_cast<uint4 *>(q) = make_uint4(out[0], out[1], out[2], out[3]);
        } else {
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                int xx = x + k;
                if (xx >= box.x0 && xx < box.x1) q[k] = (uint8_t)(out[k >> 2] >> (8 * (k & 3)));
            }
        }
        hm = hc;
        hc = hp;
        cc = cn;
    }
}

}

int sweep_gol(b200geo_grid *g, const Box& box, cudaStream_t s)
{
    const MemberLayout& L = g->m[0];
    const uint8_t *src = (const uint8_t *)g->member_ptr(0, 0) + L.origin;
    uint8_t *dst = (uint8_t *)g->member_ptr(0, 1) + L.origin;
    int xa = box.x0 & ~15;
    int groups = (box.x1 - xa + 15) / 16;
    int ny = box.y1 - box.y0;
    int gx = (groups + 127) / 128;
    int rows = 64;
    while (rows > 8 && (int64_t)gx * ((ny + rows - 1) / rows) < 148 * 96) rows /= 2;
    if (g_tuning.gol_rows > 0) rows = g_tuning.gol_rows;
    dim3 grid(gx, (ny + rows - 1) / rows);
    if (grid.y > 65535) return fail(B200GEO_ERR_OUT_OF_RANGE, "grid dimension too large");
    gol_kernel<<<grid, 128, 0, s>>>(src, dst, L.pitch, box, xa, rows, g->d[0] + g->g[0]);
    count_launch();
    return check_cuda(cudaGetLastError(), "gol sweep");
}

}
