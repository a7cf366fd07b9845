// Ghost-layer maintenance and region (de)serialisation kernels.
//
//  fill_edge    : SoAGrid::setEdge / SetContent (storage/soagrid.h:23-115, 486-494) — writes the
//                 constant edge cell into every EDGE ghost layer.
//  refresh_wrap : the periodic images a Torus axis needs (geometry/topologies.h:185-199); the
//                 reference resolves wrap-around per access (fixedneighborhoodupdatefunctor.h:185-198),
//                 here the images are materialised once per sweep so the sweep kernels stay branch-free.
//  copy_region  : SoAGrid::saveRegion / loadRegion (storage/soagrid.h:523-576); the reference's CUDA
//                 variant launches one kernel per streak (LFA detail/save_functor.hpp:115-131), this is
//                 one launch for the whole streak list. HBM-bound byte shuffling, no reuse.
#include "grid.h"

#include <vector>

namespace b200geo {

// fill a box of one member array (padded coordinates) with a constant
template<typename T>
__global__ void fill_box_kernel(T *base, int64_t pitch, int64_t plane, int x0, int y0, int z0, int w, int h, int d, T value)
{
    int64_t n = (int64_t)w * h * d;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int x = (int)(i % w);
        int64_t r = i / w;
        int y = (int)(r % h), z = (int)(r / h);
        base[(int64_t)(z0 + z) * plane + (int64_t)(y0 + y) * pitch + x0 + x] = value;
    }
}

// Only the EDGE ghost layers are touched (work ~ surface, not volume): whole padded planes below / above
// the grid, whole padded rows in front of / behind it in the remaining planes, the x ghost columns of the
// interior rows.
template<typename T>
static void launch_fill_edge(b200geo_grid *g, int m, int which, cudaStream_t s)
{
    const MemberLayout& L = g->m[m];
    T value;
    memcpy(&value, g->edge + L.edge_offset, sizeof(T));
    const int nx = g->d[0], ny = g->d[1], nz = g->d[2], gx = g->g[0], gy = g->g[1], gz = g->g[2];
    const int px = nx + 2 * gx, py = ny + 2 * gy;
    const int (*mode)[2] = g->desc.ghost_mode;
    T *base = (T *)g->member_ptr(m, which);
    auto fill = [&](int x0, int y0, int z0, int w, int h, int d) {
        int64_t n = (int64_t)w * h * d;
        if (n <= 0) return;
        int blocks = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
        fill_box_kernel<T><<<blocks, 256, 0, s>>>(base, L.pitch, L.plane, x0, y0, z0, w, h, d, value);
        count_launch();
    };
    const int x_lo = L.lead - gx;  // padded x of the first ghost column
    if (mode[2][0] == B200GEO_GHOST_EDGE) fill(x_lo, 0, 0, px, py, gz);
    if (mode[2][1] == B200GEO_GHOST_EDGE) fill(x_lo, 0, gz + nz, px, py, gz);
    if (mode[1][0] == B200GEO_GHOST_EDGE) fill(x_lo, 0, 0, px, gy, nz + 2 * gz);
    if (mode[1][1] == B200GEO_GHOST_EDGE) fill(x_lo, gy + ny, 0, px, gy, nz + 2 * gz);
    if (mode[0][0] == B200GEO_GHOST_EDGE) fill(x_lo, 0, 0, gx, py, nz + 2 * gz);
    if (mode[0][1] == B200GEO_GHOST_EDGE) fill(L.lead + nx, 0, 0, gx, py, nz + 2 * gz);
}

int fill_edge(b200geo_grid *g, int which, cudaStream_t s)
{
    for (int m = 0; m < g->n; ++m) {
        switch (g->m[m].elem) {
        case 1: launch_fill_edge<uint8_t>(g, m, which, s); break;
        case 2: launch_fill_edge<uint16_t>(g, m, which, s); break;
        case 4: launch_fill_edge<uint32_t>(g, m, which, s); break;
        default: launch_fill_edge<uint64_t>(g, m, which, s); break;
        }
    }
    return check_cuda(cudaGetLastError(), "fill_edge");
}

// copy a box of elements inside one member array (padded coordinates)
template<typename T>
__global__ void copy_box_kernel(T *base, int64_t pitch, int64_t plane,
                                int sx, int sy, int sz, int dx, int dy, int dz, int w, int h, int d)
{
    int64_t n = (int64_t)w * h * d;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int x = (int)(i % w);
        int64_t r = i / w;
        int y = (int)(r % h), z = (int)(r / h);
        base[(int64_t)(dz + z) * plane + (int64_t)(dy + y) * pitch + dx + x] =
            base[(int64_t)(sz + z) * plane + (int64_t)(sy + y) * pitch + sx + x];
    }
}

static void launch_copy_box(b200geo_grid *g, int m, int sx, int sy, int sz, int dx, int dy, int dz,
                            int w, int h, int d, cudaStream_t s)
{
    const MemberLayout& L = g->m[m];
    int64_t n = (int64_t)w * h * d;
    if (n <= 0) return;
    int blocks = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
    char *base = g->member_ptr(m, 0);
#define B200GEO_COPY_BOX(T) copy_box_kernel<T><<<blocks, 256, 0, s>>>((T *)base, L.pitch, L.plane, sx, sy, sz, dx, dy, dz, w, h, d)
    switch (L.elem) {
    case 1: B200GEO_COPY_BOX(uint8_t); break;
    case 2: B200GEO_COPY_BOX(uint16_t); break;
    case 4: B200GEO_COPY_BOX(uint32_t); break;
    default: B200GEO_COPY_BOX(uint64_t); break;
    }
#undef B200GEO_COPY_BOX
    count_launch();
}

int refresh_wrap(b200geo_grid *g, cudaStream_t s)
{
    const int (*mode)[2] = g->desc.ghost_mode;
    bool wx = mode[0][0] == B200GEO_GHOST_WRAP, wy = mode[1][0] == B200GEO_GHOST_WRAP, wz = mode[2][0] == B200GEO_GHOST_WRAP;
    if (!wx && !wy && !wz) return B200GEO_OK;
    int nx = g->d[0], ny = g->d[1], nz = g->d[2], gx = g->g[0], gy = g->g[1], gz = g->g[2];
    int pz = nz + 2 * gz;
    for (int m = 0; m < g->n; ++m) {
        const MemberLayout& L = g->m[m];
        int lead = L.lead;
        if (wx) {  // interior rows of every plane (ghost planes included: they may hold peer data);
                   // for 2-D slabs the PEER ghost ROWS hold peer data and need their x images as well
            bool peer_rows = mode[1][0] == B200GEO_GHOST_PEER || mode[1][1] == B200GEO_GHOST_PEER;
            int y0 = peer_rows ? 0 : gy, rows = peer_rows ? ny + 2 * gy : ny;
            launch_copy_box(g, m, lead + nx - gx, y0, 0, lead - gx, y0, 0, gx, rows, pz, s);
            launch_copy_box(g, m, lead, y0, 0, lead + nx, y0, 0, gx, rows, pz, s);
        }
        if (wy) {  // whole padded rows, so that the corners pick up the x images
            int w = nx + 2 * gx;
            launch_copy_box(g, m, lead - gx, gy + ny - gy, 0, lead - gx, 0, 0, w, gy, pz, s);
            launch_copy_box(g, m, lead - gx, gy, 0, lead - gx, gy + ny, 0, w, gy, pz, s);
        }
        if (wz) {  // whole padded planes
            int w = nx + 2 * gx, h = ny + 2 * gy;
            launch_copy_box(g, m, lead - gx, 0, gz + nz - gz, lead - gx, 0, 0, w, h, gz, s);
            launch_copy_box(g, m, lead - gx, 0, gz, lead - gx, 0, gz + nz, w, h, gz, s);
        }
    }
    return check_cuda(cudaGetLastError(), "refresh_wrap");
}

struct MemberTable {
    int n;
    int elem[B200GEO_MAX_MEMBERS];
    int lead[B200GEO_MAX_MEMBERS];
    int64_t pitch[B200GEO_MAX_MEMBERS];
    int64_t plane[B200GEO_MAX_MEMBERS];
    int64_t offset[B200GEO_MAX_MEMBERS];
};

// One block per streak (grid-stride); buf is member-major with `count` cells per member.
template<bool SAVE>
__global__ void copy_region_kernel(char *grid_buf, MemberTable t, int gy, int gz,
                                   const int32_t *streaks, const int64_t *prefix, int n_streaks,
                                   char *buf, int64_t count)
{
    for (int s = blockIdx.x; s < n_streaks; s += gridDim.x) {
        int x0 = streaks[4 * s], y = streaks[4 * s + 1], z = streaks[4 * s + 2], len = streaks[4 * s + 3] - x0;
        int64_t pos = prefix[s];
        int64_t boff = 0;
        for (int m = 0; m < t.n; ++m) {
            int e = t.elem[m];
            char *gp = grid_buf + t.offset[m] +
                ((int64_t)(z + gz) * t.plane[m] + (int64_t)(y + gy) * t.pitch[m] + t.lead[m] + x0) * e;
            char *bp = buf + boff + pos * e;
            // mixed member sizes can leave a member's block of the packed buffer unaligned
            if (((uintptr_t)bp % e) != 0) {
                for (int x = threadIdx.x; x < len * e; x += blockDim.x) {
                    if (SAVE) bp[x] = gp[x]; else gp[x] = bp[x];
                }
                boff += count * e;
                continue;
            }
            for (int x = threadIdx.x; x < len; x += blockDim.x) {
                if (e == 8) {
                    if (SAVE) ((uint64_t *)bp)[x] = ((const uint64_t *)gp)[x]; else ((uint64_t *)gp)[x] = ((const uint64_t *)bp)[x];
                } else if (e == 4) {
                    if (SAVE) ((uint32_t *)bp)[x] = ((const uint32_t *)gp)[x]; else ((uint32_t *)gp)[x] = ((const uint32_t *)bp)[x];
                } else if (e == 2) {
                    if (SAVE) ((uint16_t *)bp)[x] = ((const uint16_t *)gp)[x]; else ((uint16_t *)gp)[x] = ((const uint16_t *)bp)[x];
                } else {
                    if (SAVE) bp[x] = gp[x]; else gp[x] = bp[x];
                }
            }
            boff += count * e;
        }
    }
}

int copy_region(b200geo_grid *g, const int32_t *streaks, int n_streaks, char *dev_buf, int64_t count,
                bool save, int which, cudaStream_t s)
{
    size_t need = (size_t)n_streaks * (4 * sizeof(int32_t) + sizeof(int64_t));
    if (g->scratch_bytes < need) {
        if (g->scratch) { cudaStreamSynchronize(s); cudaFree(g->scratch); g->scratch = 0; g->scratch_bytes = 0; }
        B200GEO_CUDA(cudaMalloc(&g->scratch, need));
        g->scratch_bytes = need;
    }
    std::vector<int64_t> prefix(n_streaks);
    int64_t pos = 0;
    for (int i = 0; i < n_streaks; ++i) {
        prefix[i] = pos;
        pos += streaks[4 * i + 3] - streaks[4 * i];
    }
    int64_t *d_prefix = (int64_t *)g->scratch;
    int32_t *d_streaks = (int32_t *)((char *)g->scratch + (size_t)n_streaks * sizeof(int64_t));
    // pageable sources: cudaMemcpyAsync stages them before returning, so the vectors may go away
    B200GEO_CUDA(cudaMemcpyAsync(d_prefix, prefix.data(), (size_t)n_streaks * sizeof(int64_t), cudaMemcpyHostToDevice, s));
    B200GEO_CUDA(cudaMemcpyAsync(d_streaks, streaks, (size_t)n_streaks * 4 * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    MemberTable t;
    t.n = g->n;
    for (int m = 0; m < g->n; ++m) {
        t.elem[m] = g->m[m].elem;
        t.lead[m] = g->m[m].lead;
        t.pitch[m] = g->m[m].pitch;
        t.plane[m] = g->m[m].plane;
        t.offset[m] = g->m[m].offset;
    }
    int blocks = n_streaks < 148 * 8 ? n_streaks : 148 * 8;
    char *grid_buf = g->buf[g->cur ^ which];
    if (save)
        copy_region_kernel<true><<<blocks, 256, 0, s>>>(grid_buf, t, g->g[1], g->g[2], d_streaks, d_prefix, n_streaks, dev_buf, count);
    else
        copy_region_kernel<false><<<blocks, 256, 0, s>>>(grid_buf, t, g->g[1], g->g[2], d_streaks, d_prefix, n_streaks, dev_buf, count);
    count_launch();
    return check_cuda(cudaGetLastError(), "copy_region");
}

}
