/* TEST INFRASTRUCTURE — a host-memory stand-in for libb200geo.so, so that the HOST logic of the C++ façade
 * (B200Grid: combined writes, cached rows, region / member streams; B200StripedGrid: routing across slabs;
 * B200Simulator / B200StripingSimulator: event protocol, step fusing; status code -> exception mapping) runs in
 * the CPU test suite, where there is no GPU. It implements the C ABI of include/b200geo.h on plain arrays; the
 * sweeps themselves are delegated to the oracle (oracle/oracle.c) — which is exactly why this file lives under
 * tests/ and is never part of the product. Slab groups are stepped by assembling the slabs into the whole
 * space, so they exercise the façade's partition bookkeeping, not the halo schedule (that is GPU-tested).
 * Built into tests/facade/_bin/*_cpu by tests/facade/Makefile, run by tests/test_facade_cpu.py. */
#include "../../include/b200geo.h"
#include "../../oracle/oracle.h"

#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

static thread_local std::string g_error;

static int fail(int status, const std::string& msg)
{
    g_error = msg;
    return status;
}

struct b200geo_grid {
    b200geo_grid_desc desc;
    int n, d[3], g[3], elem[B200GEO_MAX_MEMBERS], cell_bytes, slab_axis, cur;
    int64_t px, py, pz;                         // padded extents
    std::vector<char> buf[2][B200GEO_MAX_MEMBERS];  // default layout: one array per member
    std::vector<char> flat[2];                      // uniform element layout: member m at stride * (bytes before m)
    char *ptr[2][B200GEO_MAX_MEMBERS];
    int64_t stride;
    int device;                                     // what the caller asked for (there is one host "device")
    unsigned char edge[8 * B200GEO_MAX_MEMBERS];
    uint64_t sweeps;

    int64_t index(int x, int y, int z) const { return ((int64_t)(z + g[2]) * py + (y + g[1])) * px + (x + g[0]); }
    int64_t cells() const { return (int64_t)d[0] * d[1] * d[2]; }
};

struct b200geo_group {
    std::vector<b200geo_grid *> g;
    bool periodic;
    uint64_t exchanges, bytes;
};

struct b200geo_boxgrid {
    b200geo_boxgrid_desc desc;
    int d[3], cap, real;
    std::vector<int32_t> counts;                // interior only, [z][y][x]
    std::vector<char> parts;                    // [z][y][x][cap][6] REAL
    bool overflow;
};

struct b200geo_boxgroup {
    std::vector<b200geo_boxgrid *> g;
    uint64_t exchanges, bytes;
};

namespace {

// dense member-major interior of the current buffer <-> grid
void gather(const b200geo_grid *g, std::vector<char>& raw)
{
    raw.resize((size_t)g->cells() * g->cell_bytes);
    size_t off = 0;
    for (int m = 0; m < g->n; ++m) {
        const int e = g->elem[m];
        for (int z = 0; z < g->d[2]; ++z)
            for (int y = 0; y < g->d[1]; ++y)
                memcpy(&raw[off + (((size_t)z * g->d[1] + y) * g->d[0]) * e], &g->ptr[g->cur][m][g->index(0, y, z) * e], (size_t)g->d[0] * e);
        off += (size_t)g->cells() * e;
    }
}

void scatter(b200geo_grid *g, int which, const std::vector<char>& raw)
{
    size_t off = 0;
    for (int m = 0; m < g->n; ++m) {
        const int e = g->elem[m];
        for (int z = 0; z < g->d[2]; ++z)
            for (int y = 0; y < g->d[1]; ++y)
                memcpy(&g->ptr[which][m][g->index(0, y, z) * e], &raw[off + (((size_t)z * g->d[1] + y) * g->d[0]) * e], (size_t)g->d[0] * e);
        off += (size_t)g->cells() * e;
    }
}

// `steps` sweeps of the kernel family over a dense member-major grid of nx * ny * nz cells
int oracle_sweeps(const b200geo_grid *like, int kernel, int nx, int ny, int nz, int steps, const std::vector<char>& in, std::vector<char>& out)
{
    out.resize(in.size());
    const bool torus = like->desc.ghost_mode[0][0] == B200GEO_GHOST_WRAP;
    switch (kernel) {
    case B200GEO_KERNEL_JACOBI6:
    case B200GEO_KERNEL_JACOBI7:
    case B200GEO_KERNEL_JACOBI27: {
        double edge;
        memcpy(&edge, like->edge, 8);
        int kind = kernel == B200GEO_KERNEL_JACOBI6 ? 6 : kernel == B200GEO_KERNEL_JACOBI7 ? 7 : 27;
        return oracle_jacobi(kind, torus, nx, ny, nz, steps, edge, (const double *)in.data(), (double *)out.data());
    }
    case B200GEO_KERNEL_GOL:
        return oracle_gol(torus, nx, ny, steps, like->edge[0] != 0, (const uint8_t *)in.data(), (uint8_t *)out.data());
    case B200GEO_KERNEL_LBM_D3Q19:
        return oracle_lbm(nx, ny, nz, steps, in.data(), out.data());
    default:
        return fail(B200GEO_ERR_LOGIC, "no kernel bound for this id");
    }
}

int region(b200geo_grid *g, const int32_t *streaks, int n, char *buf, bool save, bool both)
{
    int64_t count = 0;
    for (int i = 0; i < n; ++i) {
        const int32_t *k = streaks + 4 * i;
        if (k[3] < k[0] || k[0] < -g->g[0] || k[3] > g->d[0] + g->g[0] || k[1] < -g->g[1] || k[1] >= g->d[1] + g->g[1] ||
            k[2] < -g->g[2] || k[2] >= g->d[2] + g->g[2])
            return fail(B200GEO_ERR_INVALID, "streak outside the grid");
        count += k[3] - k[0];
    }
    int64_t moff = 0;
    for (int m = 0; m < g->n; ++m) {
        const int e = g->elem[m];
        int64_t pos = 0;
        for (int i = 0; i < n; ++i) {
            const int32_t *k = streaks + 4 * i;
            const int64_t len = k[3] - k[0], at = g->index(k[0], k[1], k[2]) * e;
            char *b = buf + moff + pos * e;
            if (save) {
                memcpy(b, &g->ptr[g->cur][m][at], len * e);
            } else {
                memcpy(&g->ptr[g->cur][m][at], b, len * e);
                if (both) memcpy(&g->ptr[g->cur ^ 1][m][at], b, len * e);
            }
            pos += len;
        }
        moff += count * e;
    }
    return B200GEO_OK;
}

int member_box(b200geo_grid *g, int m, const int32_t o[3], const int32_t d[3], char *dense, bool load, bool both)
{
    if (m < 0 || m >= g->n) return fail(B200GEO_ERR_INVALID, "bad member index");
    for (int i = 0; i < 3; ++i)
        if (d[i] < 0 || o[i] < -g->g[i] || o[i] + d[i] > g->d[i] + g->g[i]) return fail(B200GEO_ERR_INVALID, "box outside the grid");
    const int e = g->elem[m];
    for (int z = 0; z < d[2]; ++z)
        for (int y = 0; y < d[1]; ++y) {
            char *row = dense + (((size_t)z * d[1] + y) * d[0]) * e;
            const int64_t at = g->index(o[0], o[1] + y, o[2] + z) * e;
            if (load) {
                memcpy(&g->ptr[g->cur][m][at], row, (size_t)d[0] * e);
                if (both) memcpy(&g->ptr[g->cur ^ 1][m][at], row, (size_t)d[0] * e);
            } else {
                memcpy(row, &g->ptr[g->cur][m][at], (size_t)d[0] * e);
            }
        }
    return B200GEO_OK;
}

size_t box_cell_bytes(const b200geo_boxgrid *g) { return (size_t)g->cap * 6 * g->real; }

}

extern "C" {

const char *b200geo_version(void) { return "b200geo mock (host memory, tests only)"; }
const char *b200geo_last_error(void) { return g_error.c_str(); }
int b200geo_device_count(void) { return 1; }
int b200geo_set_tuning(const char *, int) { return B200GEO_OK; }
uint64_t b200geo_launch_count(void) { return 0; }

static int create_grid(const b200geo_grid_desc *desc, int64_t stride, b200geo_grid **out);

int b200geo_grid_create(const b200geo_grid_desc *desc, int device, b200geo_grid **out)
{
    int rc = create_grid(desc, 0, out);
    if (rc == B200GEO_OK) (*out)->device = device;
    return rc;
}

int b200geo_grid_uniform_min_stride(const b200geo_grid_desc *desc, int64_t *min_stride)
{
    int64_t elems = 1;
    for (int i = 0; i < 3; ++i) elems *= desc->dim[i] + 2 * desc->ghost[i];
    *min_stride = (elems + 255) / 256 * 256;
    return B200GEO_OK;
}

int b200geo_grid_create_uniform(const b200geo_grid_desc *desc, int device, int64_t member_stride, b200geo_grid **out)
{
    int64_t least = 0;
    b200geo_grid_uniform_min_stride(desc, &least);
    if (member_stride <= 0 || member_stride % 256 != 0) return fail(B200GEO_ERR_INVALID, "member stride must be a positive multiple of 256 elements");
    if (member_stride < least) return fail(B200GEO_ERR_INVALID, "member stride smaller than the padded grid");
    int rc = create_grid(desc, member_stride, out);
    if (rc == B200GEO_OK) (*out)->device = device;
    return rc;
}

int b200geo_grid_member_stride(const b200geo_grid *g, int64_t *member_stride)
{
    *member_stride = g->stride;
    return B200GEO_OK;
}

static int create_grid(const b200geo_grid_desc *desc, int64_t stride, b200geo_grid **out)
{
    if (!desc || !out) return fail(B200GEO_ERR_INVALID, "null argument");
    if (desc->n_members < 1 || desc->n_members > B200GEO_MAX_MEMBERS) return fail(B200GEO_ERR_INVALID, "n_members out of range");
    int slab = (desc->dim[2] == 1 && desc->ghost[2] == 0) ? 1 : 2;
    for (int i = 0; i < 3; ++i) {
        if (desc->dim[i] < 1) return fail(B200GEO_ERR_INVALID, "grid dimension must be >= 1");
        if (desc->ghost[i] < 0 || desc->ghost[i] > (i == 0 ? 16 : 65536)) return fail(B200GEO_ERR_INVALID, "ghost width out of range");
        for (int s = 0; s < 2; ++s) {
            int mode = desc->ghost_mode[i][s];
            if (mode < B200GEO_GHOST_EDGE || mode > B200GEO_GHOST_PEER) return fail(B200GEO_ERR_INVALID, "bad ghost mode");
            if (mode == B200GEO_GHOST_PEER && i != slab)
                return fail(B200GEO_ERR_LOGIC, "PEER ghost layers are supported on the last axis only (slab partition)");
            if (mode == B200GEO_GHOST_WRAP && desc->ghost[i] > desc->dim[i]) return fail(B200GEO_ERR_INVALID, "wrap ghost wider than the grid");
        }
    }
    b200geo_grid *g = new b200geo_grid();
    g->desc = *desc;
    g->n = desc->n_members;
    g->cell_bytes = 0;
    g->cur = 0;
    g->sweeps = 0;
    g->slab_axis = slab;
    memset(g->edge, 0, sizeof(g->edge));
    for (int i = 0; i < 3; ++i) {
        g->d[i] = desc->dim[i];
        g->g[i] = desc->ghost[i];
    }
    g->px = g->d[0] + 2 * g->g[0];
    g->py = g->d[1] + 2 * g->g[1];
    g->pz = g->d[2] + 2 * g->g[2];
    for (int m = 0; m < g->n; ++m) {
        int e = desc->member_bytes[m];
        if (e != 1 && e != 2 && e != 4 && e != 8) {
            delete g;
            return fail(B200GEO_ERR_INVALID, "member size must be 1, 2, 4 or 8 bytes");
        }
        g->elem[m] = e;
        g->cell_bytes += e;
        if (stride == 0) {
            for (int b = 0; b < 2; ++b) {
                g->buf[b][m].assign((size_t)(g->px * g->py * g->pz) * e, 0);
                g->ptr[b][m] = g->buf[b][m].data();
            }
        }
    }
    g->stride = stride;
    if (stride > 0) {
        for (int b = 0; b < 2; ++b) {
            g->flat[b].assign((size_t)stride * g->cell_bytes, 0);
            size_t before = 0;
            for (int m = 0; m < g->n; ++m) {
                g->ptr[b][m] = g->flat[b].data() + (size_t)stride * before;
                before += g->elem[m];
            }
        }
    }
    *out = g;
    return B200GEO_OK;
}

int b200geo_grid_destroy(b200geo_grid *g) { delete g; return B200GEO_OK; }
int b200geo_grid_buffer_bytes(const b200geo_grid *g, uint64_t *bytes) { *bytes = (uint64_t)(g->px * g->py * g->pz) * g->cell_bytes; return B200GEO_OK; }
int b200geo_grid_device(const b200geo_grid *g, int *device) { *device = g->device; return B200GEO_OK; }

int b200geo_grid_layout(const b200geo_grid *g, int, int64_t *pitch, int64_t *plane, int64_t *origin)
{
    if (pitch) *pitch = g->px;
    if (plane) *plane = g->px * g->py;
    if (origin) *origin = g->index(0, 0, 0);
    return B200GEO_OK;
}

int b200geo_grid_member_ptr(const b200geo_grid *g, int m, int which, void **ptr)
{
    *ptr = g->ptr[g->cur ^ which][m];
    return B200GEO_OK;
}

int b200geo_grid_set_edge(b200geo_grid *g, const void *cell, void *)
{
    memcpy(g->edge, cell, g->cell_bytes);
    int off = 0;
    for (int m = 0; m < g->n; ++m) {
        const int e = g->elem[m];
        for (int z = -g->g[2]; z < g->d[2] + g->g[2]; ++z)
            for (int y = -g->g[1]; y < g->d[1] + g->g[1]; ++y)
                for (int x = -g->g[0]; x < g->d[0] + g->g[0]; ++x) {
                    const int c[3] = {x, y, z};
                    bool is_edge = false;
                    for (int i = 0; i < 3; ++i) {
                        if (c[i] < 0 && g->desc.ghost_mode[i][0] == B200GEO_GHOST_EDGE) is_edge = true;
                        if (c[i] >= g->d[i] && g->desc.ghost_mode[i][1] == B200GEO_GHOST_EDGE) is_edge = true;
                    }
                    if (!is_edge) continue;
                    for (int b = 0; b < 2; ++b) memcpy(&g->ptr[b][m][g->index(x, y, z) * e], g->edge + off, e);
                }
        off += e;
    }
    return B200GEO_OK;
}

int b200geo_grid_get_edge(const b200geo_grid *g, void *cell) { memcpy(cell, g->edge, g->cell_bytes); return B200GEO_OK; }

int b200geo_grid_load_member(b200geo_grid *g, int m, const int32_t o[3], const int32_t d[3], const void *src, int, int both, void *)
{
    return member_box(g, m, o, d, (char *)const_cast<void *>(src), true, both != 0);
}

int b200geo_grid_save_member(const b200geo_grid *g, int m, const int32_t o[3], const int32_t d[3], void *dst, int, void *)
{
    return member_box(const_cast<b200geo_grid *>(g), m, o, d, (char *)dst, false, false);
}

int b200geo_grid_load_region(b200geo_grid *g, const int32_t *streaks, int n, const void *buf, int, int both, void *)
{
    if (n <= 0) return B200GEO_OK;
    return region(g, streaks, n, (char *)const_cast<void *>(buf), false, both != 0);
}

int b200geo_grid_save_region(const b200geo_grid *g, const int32_t *streaks, int n, void *buf, int, void *)
{
    if (n <= 0) return B200GEO_OK;
    return region(const_cast<b200geo_grid *>(g), streaks, n, (char *)buf, true, false);
}

int b200geo_step(b200geo_grid *g, int kernel, const void *, uint32_t, uint32_t n_steps, void *)
{
    for (int i = 0; i < 3; ++i)
        for (int s = 0; s < 2; ++s)
            if (g->desc.ghost_mode[i][s] == B200GEO_GHOST_PEER) return fail(B200GEO_ERR_LOGIC, "mock: slabs are stepped through their group");
    if (n_steps == 0) return B200GEO_OK;
    std::vector<char> in, out;
    gather(g, in);
    int rc = oracle_sweeps(g, kernel, g->d[0], g->d[1], g->d[2], (int)n_steps, in, out);
    if (rc) return rc < 0 ? rc : fail(B200GEO_ERR_INVALID, "oracle failed");
    // an odd number of sweeps ends in the other buffer, like the real engine
    int target = (n_steps & 1) ? g->cur ^ 1 : g->cur;
    scatter(g, target, out);
    g->cur = target;
    g->sweeps += n_steps;
    return B200GEO_OK;
}

// n sweeps over the box [o, o + d): current buffer -> scratch buffer, no swap. The oracle runs on the box plus a ring of
// n cells (ghost cells included: they are real cells of the padded arrays), only the box is written back.
int b200geo_update_box_n(b200geo_grid *g, int kernel, const void *, uint32_t, const int32_t *o, const int32_t *d, uint32_t n_sweeps, void *)
{
    const int n = (int)n_sweeps;
    if (n < 1) return fail(B200GEO_ERR_INVALID, "n_sweeps must be >= 1");
    if (d[0] <= 0 || d[1] <= 0 || d[2] <= 0) return B200GEO_OK;
    int lo[3], hi[3];
    for (int i = 0; i < 3; ++i) {
        const bool used = g->d[i] > 1 || g->g[i] > 0;
        lo[i] = o[i] - (used ? n : 0);
        hi[i] = o[i] + d[i] + (used ? n : 0);
        // beyond a Cube boundary there is nothing but the constant edge cell: one ring of it is all a sweep reads
        if (lo[i] < -g->g[i] && g->desc.ghost_mode[i][0] == B200GEO_GHOST_EDGE) lo[i] = -g->g[i];
        if (hi[i] > g->d[i] + g->g[i] && g->desc.ghost_mode[i][1] == B200GEO_GHOST_EDGE) hi[i] = g->d[i] + g->g[i];
        if (lo[i] < -g->g[i] || hi[i] > g->d[i] + g->g[i]) return fail(B200GEO_ERR_INVALID, "box outside the updatable area");
    }
    const int ex = hi[0] - lo[0], ey = hi[1] - lo[1], ez = hi[2] - lo[2];
    const size_t cells = (size_t)ex * ey * ez;
    std::vector<char> in(cells * g->cell_bytes), out;
    size_t off = 0;
    for (int m = 0; m < g->n; ++m) {
        const int e = g->elem[m];
        for (int z = 0; z < ez; ++z)
            for (int y = 0; y < ey; ++y)
                memcpy(&in[off + (((size_t)z * ey + y) * ex) * e], &g->ptr[g->cur][m][g->index(lo[0], lo[1] + y, lo[2] + z) * e], (size_t)ex * e);
        off += cells * e;
    }
    // the window's own boundary is never a topological one here: periodic images are cells of the ghost ring
    b200geo_grid cube = *g;
    for (int i = 0; i < 3; ++i) cube.desc.ghost_mode[i][0] = cube.desc.ghost_mode[i][1] = B200GEO_GHOST_EDGE;
    for (int sweep = 0; sweep < n; ++sweep) {
        int rc = oracle_sweeps(&cube, kernel, ex, ey, ez, 1, in, out);
        if (rc) return rc < 0 ? rc : fail(B200GEO_ERR_INVALID, "oracle failed");
        // cells of the constant edge ring (EDGE sides) inside the window never change
        off = 0;
        for (int m = 0; m < g->n; ++m) {
            const int e = g->elem[m];
            for (int z = 0; z < ez; ++z)
                for (int y = 0; y < ey; ++y)
                    for (int x = 0; x < ex; ++x) {
                        const int c[3] = {lo[0] + x, lo[1] + y, lo[2] + z};
                        bool is_edge = false;
                        for (int i = 0; i < 3; ++i) {
                            if (c[i] < 0 && g->desc.ghost_mode[i][0] == B200GEO_GHOST_EDGE) is_edge = true;
                            if (c[i] >= g->d[i] && g->desc.ghost_mode[i][1] == B200GEO_GHOST_EDGE) is_edge = true;
                        }
                        if (is_edge) {
                            const size_t at = off + (((size_t)z * ey + y) * ex + x) * e;
                            memcpy(&out[at], &in[at], e);
                        }
                    }
            off += cells * e;
        }
        if (sweep + 1 < n) in.swap(out);
    }
    off = 0;
    for (int m = 0; m < g->n; ++m) {
        const int e = g->elem[m];
        for (int z = 0; z < d[2]; ++z)
            for (int y = 0; y < d[1]; ++y)
                memcpy(&g->ptr[g->cur ^ 1][m][g->index(o[0], o[1] + y, o[2] + z) * e],
                       &out[off + (((size_t)(z + o[2] - lo[2]) * ey + (y + o[1] - lo[1])) * ex + (o[0] - lo[0])) * e], (size_t)d[0] * e);
        off += cells * e;
    }
    return B200GEO_OK;
}

int b200geo_update_box(b200geo_grid *g, int kernel, const void *params, uint32_t nano, const int32_t *o, const int32_t *d, void *stream)
{
    return b200geo_update_box_n(g, kernel, params, nano, o, d, 1, stream);
}
int b200geo_swap(b200geo_grid *g) { g->cur ^= 1; return B200GEO_OK; }
// periodic images of WRAP axes in the current buffer, axis by axis (x first, so that corners end up right)
int b200geo_refresh_ghosts(b200geo_grid *g, void *)
{
    for (int axis = 0; axis < 3; ++axis) {
        if (g->desc.ghost_mode[axis][0] != B200GEO_GHOST_WRAP || g->g[axis] == 0) continue;
        for (int m = 0; m < g->n; ++m) {
            const int e = g->elem[m];
            char *base = g->ptr[g->cur][m];
            for (int z = -g->g[2]; z < g->d[2] + g->g[2]; ++z)
                for (int y = -g->g[1]; y < g->d[1] + g->g[1]; ++y)
                    for (int x = -g->g[0]; x < g->d[0] + g->g[0]; ++x) {
                        int c[3] = {x, y, z};
                        if (c[axis] >= 0 && c[axis] < g->d[axis]) continue;
                        int s[3] = {x, y, z};
                        s[axis] = (c[axis] + g->d[axis]) % g->d[axis];
                        memcpy(base + g->index(x, y, z) * e, base + g->index(s[0], s[1], s[2]) * e, e);
                    }
        }
    }
    return B200GEO_OK;
}
int b200geo_sync(void *) { return B200GEO_OK; }
int b200geo_grid_sync(const b200geo_grid *, void *) { return B200GEO_OK; }
/* the mock engine executes every call synchronously: streams are tokens, waiting is a no-op */
static int g_mock_streams = 0;
int b200geo_stream_create(int, void **stream) { *stream = (void *)(intptr_t)(0x1000 + ++g_mock_streams); return B200GEO_OK; }
int b200geo_stream_destroy(int, void *) { return B200GEO_OK; }
int b200geo_stream_wait(int, void *, void *) { return B200GEO_OK; }
int b200geo_device_alloc(int, uint64_t bytes, void **ptr) { *ptr = malloc(bytes ? (size_t)bytes : 1); return *ptr ? B200GEO_OK : fail(B200GEO_ERR_NOMEM, "out of memory"); }
int b200geo_device_free(int, void *ptr) { free(ptr); return B200GEO_OK; }
int b200geo_device_copy(int, void *dst, int, const void *src, uint64_t bytes) { memcpy(dst, src, (size_t)bytes); return B200GEO_OK; }
int b200geo_host_alloc(uint64_t bytes, void **ptr) { *ptr = malloc(bytes ? (size_t)bytes : 1); return *ptr ? B200GEO_OK : fail(B200GEO_ERR_NOMEM, "out of memory"); }
int b200geo_host_free(void *ptr) { free(ptr); return B200GEO_OK; }
int b200geo_halo_block(const b200geo_grid *, int, int, int, int, void **, uint64_t *) { return fail(B200GEO_ERR_LOGIC, "not in the mock"); }
int b200geo_halo_block_in(const b200geo_grid *, int, int, int, int, int, void **, uint64_t *) { return fail(B200GEO_ERR_LOGIC, "not in the mock"); }
int b200geo_grid_ipc_export(const b200geo_grid *, int, void *) { return fail(B200GEO_ERR_LOGIC, "not in the mock"); }
int b200geo_grid_ipc_open(b200geo_grid *, int, int, const void *) { return fail(B200GEO_ERR_LOGIC, "not in the mock"); }
int b200geo_halo_push(b200geo_grid *, int, int, void *) { return fail(B200GEO_ERR_LOGIC, "not in the mock"); }
int b200geo_halo_mark_valid(b200geo_grid *, int, int) { return B200GEO_OK; }
int b200geo_stats_enable(b200geo_grid *, int) { return B200GEO_OK; }
int b200geo_stats(b200geo_grid *g, double out[3]) { out[0] = out[1] = 0; out[2] = (double)g->sweeps; return B200GEO_OK; }

/* ---- slab groups: assembled into the whole space ---------------------------------------------------------- */
int b200geo_group_create(b200geo_grid *const *grids, int n, int periodic, b200geo_group **out)
{
    if (!grids || !out || n < 1 || n > 16) return fail(B200GEO_ERR_INVALID, "bad slab group");
    for (int i = 0; i < n; ++i) {
        const b200geo_grid *g = grids[i];
        const int a = g->slab_axis;
        if (n > 1) {
            bool low = i > 0 || periodic, high = i < n - 1 || periodic;
            if ((g->desc.ghost_mode[a][0] == B200GEO_GHOST_PEER) != low || (g->desc.ghost_mode[a][1] == B200GEO_GHOST_PEER) != high)
                return fail(B200GEO_ERR_INVALID, "slab faces towards a neighbour must be PEER ghost layers, outer faces must not");
            if (g->d[a] < g->g[a]) return fail(B200GEO_ERR_INVALID, "slab thinner than the ghost zone");
        }
    }
    b200geo_group *grp = new b200geo_group();
    grp->g.assign(grids, grids + n);
    grp->periodic = periodic != 0;
    grp->exchanges = grp->bytes = 0;
    *out = grp;
    return B200GEO_OK;
}

int b200geo_group_destroy(b200geo_group *grp) { delete grp; return B200GEO_OK; }
int b200geo_group_invalidate(b200geo_group *) { return B200GEO_OK; }
int b200geo_group_exchange(b200geo_group *) { return B200GEO_OK; }

int b200geo_group_step(b200geo_group *grp, int kernel, const void *params, uint32_t first, uint32_t n_steps)
{
    if (grp->g.size() == 1) return b200geo_step(grp->g[0], kernel, params, first, n_steps, 0);
    if (n_steps == 0) return B200GEO_OK;
    b200geo_grid *g0 = grp->g[0];
    const int a = g0->slab_axis;
    int dims[3] = {g0->d[0], g0->d[1], g0->d[2]};
    dims[a] = 0;
    for (size_t s = 0; s < grp->g.size(); ++s) dims[a] += grp->g[s]->d[a];
    const int64_t cells = (int64_t)dims[0] * dims[1] * dims[2];
    std::vector<char> in((size_t)cells * g0->cell_bytes), out, part;
    // slabs are contiguous along the slowest used axis: member by member, slab after slab
    size_t moff = 0;
    for (int m = 0; m < g0->n; ++m) {
        size_t pos = 0;
        for (size_t s = 0; s < grp->g.size(); ++s) {
            b200geo_grid *g = grp->g[s];
            gather(g, part);
            size_t poff = 0;
            for (int k = 0; k < m; ++k) poff += (size_t)g->cells() * g->elem[k];
            memcpy(&in[moff + pos], &part[poff], (size_t)g->cells() * g->elem[m]);
            pos += (size_t)g->cells() * g->elem[m];
        }
        moff += (size_t)cells * g0->elem[m];
    }
    // a Torus group: the wrap flag lives in the x axis of the slabs (the slab axis itself is PEER)
    int rc = oracle_sweeps(g0, kernel, dims[0], dims[1], dims[2], (int)n_steps, in, out);
    if (rc) return rc < 0 ? rc : fail(B200GEO_ERR_INVALID, "oracle failed");
    moff = 0;
    std::vector<std::vector<char> > parts(grp->g.size());
    for (size_t s = 0; s < grp->g.size(); ++s) parts[s].resize((size_t)grp->g[s]->cells() * g0->cell_bytes);
    for (int m = 0; m < g0->n; ++m) {
        size_t pos = 0;
        for (size_t s = 0; s < grp->g.size(); ++s) {
            b200geo_grid *g = grp->g[s];
            size_t poff = 0;
            for (int k = 0; k < m; ++k) poff += (size_t)g->cells() * g->elem[k];
            memcpy(&parts[s][poff], &out[moff + pos], (size_t)g->cells() * g->elem[m]);
            pos += (size_t)g->cells() * g->elem[m];
        }
        moff += (size_t)cells * g0->elem[m];
    }
    for (size_t s = 0; s < grp->g.size(); ++s) {
        b200geo_grid *g = grp->g[s];
        int target = (n_steps & 1) ? g->cur ^ 1 : g->cur;
        scatter(g, target, parts[s]);
        g->cur = target;
        g->sweeps += n_steps;
    }
    const int w = g0->g[a] > 0 ? g0->g[a] : 1;
    grp->exchanges += (n_steps + w - 1) / w;
    grp->bytes += 1;
    return B200GEO_OK;
}

// Callback-driven stepping (generic device paths): per sweep every PEER ghost slice is filled from the neighbour's
// outermost owned slices (whole padded slices, x / y ghosts included — what the real group copies over NVLink), the
// periodic images of the other axes are refreshed, the callback updates each slab's whole box, the buffers swap. One
// exchange per sweep, no rim / interior split: this checks the callback plumbing and the ghost geometry, not the schedule.
int b200geo_group_step_with(b200geo_group *grp, b200geo_update_fn update, void *ctx, uint32_t first, uint32_t n_steps)
{
    const int n = (int)grp->g.size();
    for (uint32_t t = 0; t < n_steps; ++t) {
        for (int s = 0; s < n && n > 1; ++s) {
            b200geo_grid *g = grp->g[s];
            const int a = g->slab_axis, w = g->g[a];
            for (int side = 0; side < 2; ++side) {
                if (g->desc.ghost_mode[a][side] != B200GEO_GHOST_PEER) continue;
                b200geo_grid *nb = grp->g[(s + (side ? 1 : n - 1)) % n];
                for (int m = 0; m < g->n; ++m) {
                    const int e = g->elem[m];
                    for (int k = 0; k < w; ++k) {
                        // ghost slice k below / above this slab <- the neighbour's k-th slice from its far / near face
                        const int dst = side ? g->d[a] + k : -w + k;
                        const int src = side ? k : nb->d[a] - w + k;
                        for (int y = (a == 1 ? 0 : -g->g[1]); y < (a == 1 ? 1 : g->d[1] + g->g[1]); ++y) {
                            const int64_t to = a == 2 ? g->index(-g->g[0], y, dst) : g->index(-g->g[0], dst, 0);
                            const int64_t from = a == 2 ? nb->index(-nb->g[0], y, src) : nb->index(-nb->g[0], src, 0);
                            memcpy(g->ptr[g->cur][m] + to * e, nb->ptr[nb->cur][m] + from * e, (size_t)g->px * e);
                        }
                    }
                }
            }
            grp->bytes += 1;
        }
        if (n > 1) ++grp->exchanges;
        for (int s = 0; s < n; ++s) {
            b200geo_grid *g = grp->g[s];
            b200geo_refresh_ghosts(g, 0);
            const int32_t origin[3] = {0, 0, 0}, dim[3] = {g->d[0], g->d[1], g->d[2]};
            int rc = update(ctx, g, first + t, origin, dim, 0);
            if (rc < 0) return rc;
        }
        for (int s = 0; s < n; ++s) {
            grp->g[s]->cur ^= 1;
            ++grp->g[s]->sweeps;
        }
    }
    return B200GEO_OK;
}
int b200geo_group_sync(b200geo_group *) { return B200GEO_OK; }
int b200geo_group_stats(const b200geo_group *grp, uint64_t out[2]) { out[0] = grp->exchanges; out[1] = grp->bytes; return B200GEO_OK; }

/* ---- container grids ---------------------------------------------------------------------------------------- */
int b200geo_boxgrid_create(const b200geo_boxgrid_desc *desc, int, b200geo_boxgrid **out)
{
    if (!desc || !out) return fail(B200GEO_ERR_INVALID, "null argument");
    if (desc->capacity < 1 || desc->capacity > 64 || (desc->real_bytes != 4 && desc->real_bytes != 8))
        return fail(B200GEO_ERR_INVALID, "bad capacity / particle type");
    b200geo_boxgrid *g = new b200geo_boxgrid();
    g->desc = *desc;
    for (int i = 0; i < 3; ++i) g->d[i] = desc->dim[i];
    g->cap = desc->capacity;
    g->real = desc->real_bytes;
    g->overflow = false;
    g->counts.assign((size_t)g->d[0] * g->d[1] * g->d[2], 0);
    g->parts.assign(g->counts.size() * box_cell_bytes(g), 0);
    *out = g;
    return B200GEO_OK;
}

int b200geo_boxgrid_destroy(b200geo_boxgrid *g) { delete g; return B200GEO_OK; }

static int box_io(b200geo_boxgrid *g, const int32_t o[3], const int32_t d[3], int32_t *counts, char *parts, bool load)
{
    for (int i = 0; i < 3; ++i)
        if (d[i] < 0 || o[i] < 0 || o[i] + d[i] > g->d[i]) return fail(B200GEO_ERR_INVALID, "mock: container box outside the interior");
    const size_t cb = box_cell_bytes(g);
    for (int z = 0; z < d[2]; ++z)
        for (int y = 0; y < d[1]; ++y)
            for (int x = 0; x < d[0]; ++x) {
                size_t dense = ((size_t)z * d[1] + y) * d[0] + x;
                size_t mine = ((size_t)(o[2] + z) * g->d[1] + (o[1] + y)) * g->d[0] + (o[0] + x);
                if (load) {
                    g->counts[mine] = counts[dense];
                    memcpy(&g->parts[mine * cb], parts + dense * cb, cb);
                } else {
                    counts[dense] = g->counts[mine];
                    memcpy(parts + dense * cb, &g->parts[mine * cb], cb);
                }
            }
    return B200GEO_OK;
}

int b200geo_boxgrid_load(b200geo_boxgrid *g, const int32_t o[3], const int32_t d[3], const int32_t *counts, const void *parts, int, int, void *)
{
    return box_io(g, o, d, const_cast<int32_t *>(counts), (char *)const_cast<void *>(parts), true);
}

int b200geo_boxgrid_save(const b200geo_boxgrid *g, const int32_t o[3], const int32_t d[3], int32_t *counts, void *parts, int, void *)
{
    return box_io(const_cast<b200geo_boxgrid *>(g), o, d, counts, (char *)parts, false);
}

static int nbody_sweeps(b200geo_boxgrid *like, int nx, int ny, int nz, const b200geo_nbody_params *p, uint32_t steps,
                        std::vector<int32_t>& counts, std::vector<char>& parts)
{
    if (!(p->cutoff <= like->desc.cell_edge)) return fail(B200GEO_ERR_INVALID, "cutoff larger than the container edge: interactions would be missed");
    std::vector<int32_t> co(counts.size());
    std::vector<char> po(parts.size());
    int rc = oracle_nbody(like->real, nx, ny, nz, like->cap, (int)steps, p->dt, p->cutoff, like->desc.cell_edge, counts.data(), parts.data(), co.data(), po.data());
    if (rc == -3) {
        like->overflow = true;
        return B200GEO_OK;   // reported by b200geo_boxgrid_check, like the device flag
    }
    if (rc) return fail(B200GEO_ERR_INVALID, "oracle failed");
    counts.swap(co);
    parts.swap(po);
    return B200GEO_OK;
}

int b200geo_boxgrid_step(b200geo_boxgrid *g, const b200geo_nbody_params *p, uint32_t, uint32_t n_steps, void *)
{
    if (g->desc.ghost_mode[2][0] == B200GEO_GHOST_PEER || g->desc.ghost_mode[2][1] == B200GEO_GHOST_PEER)
        return fail(B200GEO_ERR_LOGIC, "mock: slabs are stepped through their group");
    if (n_steps == 0) return B200GEO_OK;
    return nbody_sweeps(g, g->d[0], g->d[1], g->d[2], p, n_steps, g->counts, g->parts);
}

int b200geo_boxgrid_check(b200geo_boxgrid *g, void *)
{
    if (g->overflow) {
        g->overflow = false;
        return fail(B200GEO_ERR_OUT_OF_RANGE, "capacity exceeded");
    }
    return B200GEO_OK;
}

int b200geo_boxgrid_halo_block(const b200geo_boxgrid *, int, int, int, int, void **, uint64_t *) { return fail(B200GEO_ERR_LOGIC, "not in the mock"); }
int b200geo_boxgrid_halo_mark_valid(b200geo_boxgrid *, int, int) { return B200GEO_OK; }

int b200geo_boxgroup_create(b200geo_boxgrid *const *grids, int n, b200geo_boxgroup **out)
{
    if (!grids || !out || n < 1 || n > 16) return fail(B200GEO_ERR_INVALID, "bad slab group");
    b200geo_boxgroup *grp = new b200geo_boxgroup();
    grp->g.assign(grids, grids + n);
    grp->exchanges = grp->bytes = 0;
    *out = grp;
    return B200GEO_OK;
}

int b200geo_boxgroup_destroy(b200geo_boxgroup *grp) { delete grp; return B200GEO_OK; }

int b200geo_boxgroup_step(b200geo_boxgroup *grp, const b200geo_nbody_params *p, uint32_t, uint32_t n_steps)
{
    if (n_steps == 0) return B200GEO_OK;
    b200geo_boxgrid *g0 = grp->g[0];
    std::vector<int32_t> counts;
    std::vector<char> parts;
    int nz = 0;
    for (size_t s = 0; s < grp->g.size(); ++s) {   // slabs along z are contiguous blocks of the dense arrays
        b200geo_boxgrid *g = grp->g[s];
        counts.insert(counts.end(), g->counts.begin(), g->counts.end());
        parts.insert(parts.end(), g->parts.begin(), g->parts.end());
        nz += g->d[2];
    }
    // the assembled space starts at container (0, 0, 0), which is what the oracle assumes
    int rc = nbody_sweeps(g0, g0->d[0], g0->d[1], nz, p, n_steps, counts, parts);
    if (rc) return rc;
    size_t c = 0;
    for (size_t s = 0; s < grp->g.size(); ++s) {
        b200geo_boxgrid *g = grp->g[s];
        std::copy(counts.begin() + c, counts.begin() + c + g->counts.size(), g->counts.begin());
        memcpy(g->parts.data(), &parts[c * box_cell_bytes(g0)], g->parts.size());
        c += g->counts.size();
        if (g0->overflow) g->overflow = true;
    }
    if (grp->g.size() > 1) grp->exchanges += n_steps;
    return B200GEO_OK;
}

int b200geo_boxgroup_sync(b200geo_boxgroup *) { return B200GEO_OK; }
int b200geo_boxgroup_stats(const b200geo_boxgroup *grp, uint64_t out[2]) { out[0] = grp->exchanges; out[1] = grp->bytes; return B200GEO_OK; }

}

/* ---- ContainerCell grids (ID-keyed cargo) ------------------------------------------------------------------- */
struct b200geo_containergrid {
    b200geo_containergrid_desc desc;
    int d[3], cap, maxnb;
    size_t cells;                       // interior containers; record `cells` is the edge container
    std::vector<int32_t> counts, ids, nbc, nbids;
    std::vector<double> values, influx;
    bool dirty;
    uint64_t rebuilds, sweeps, links;
};

template<typename T>
static void container_array(std::vector<T>& store, T *user, const std::vector<int32_t>& counts, int per, int per_slot, size_t first,
                            const int32_t o[3], const int32_t d[3], const int gd[3], bool edge, bool load)
{
    if (!user) return;
    size_t b = 0;
    for (int z = 0; z < d[2]; ++z) {
        for (int y = 0; y < d[1]; ++y) {
            for (int x = 0; x < d[0]; ++x, ++b) {
                size_t c = edge ? first : ((size_t)(z + o[2]) * gd[1] + (y + o[1])) * gd[0] + (x + o[0]);
                for (int e = 0; e < per; ++e) {
                    if (load) store[c * per + e] = user[b * per + e];
                    else user[b * per + e] = (per_slot > 0 && e / per_slot >= counts[c]) ? T(0) : store[c * per + e];
                }
            }
        }
    }
}

extern "C" {

int b200geo_containergrid_create(const b200geo_containergrid_desc *desc, int, b200geo_containergrid **out)
{
    if (!desc || !out) return fail(B200GEO_ERR_INVALID, "null argument");
    if ((desc->n_dims != 2 && desc->n_dims != 3) || (desc->n_dims == 2 && desc->dim[2] != 1))
        return fail(B200GEO_ERR_INVALID, "n_dims must be 2 or 3");
    if (desc->capacity < 1 || desc->capacity > 4096 || desc->max_neighbors < 1 || desc->max_neighbors > 64)
        return fail(B200GEO_ERR_INVALID, "bad capacity / max_neighbors");
    for (int i = 0; i < 3; ++i) {
        if (desc->dim[i] < 1) return fail(B200GEO_ERR_INVALID, "grid dimension must be >= 1");
        if (desc->ghost_mode[i][0] != desc->ghost_mode[i][1] || desc->ghost_mode[i][0] == B200GEO_GHOST_PEER)
            return fail(B200GEO_ERR_INVALID, "ghost mode must be EDGE or WRAP, alike on both sides of an axis");
    }
    b200geo_containergrid *g = new b200geo_containergrid();
    g->desc = *desc;
    for (int i = 0; i < 3; ++i) g->d[i] = desc->dim[i];
    g->cap = desc->capacity;
    g->maxnb = desc->max_neighbors;
    g->cells = (size_t)g->d[0] * g->d[1] * g->d[2];
    size_t slots = (g->cells + 1) * g->cap;
    g->counts.assign(g->cells + 1, 0);
    g->ids.assign(slots, 0);
    g->nbc.assign(slots, 0);
    g->nbids.assign(slots * g->maxnb, 0);
    g->values.assign(slots, 0);
    g->influx.assign(slots, 0);
    g->dirty = true;
    g->rebuilds = g->sweeps = g->links = 0;
    *out = g;
    return B200GEO_OK;
}

int b200geo_containergrid_destroy(b200geo_containergrid *g) { delete g; return B200GEO_OK; }

static int container_io(b200geo_containergrid *g, const int32_t o[3], const int32_t d[3], const b200geo_container_box *box, bool edge, bool load)
{
    container_array<int32_t>(g->counts, box->counts, g->counts, 1, 0, g->cells, o, d, g->d, edge, load);
    container_array<int32_t>(g->ids, box->ids, g->counts, g->cap, 1, g->cells, o, d, g->d, edge, load);
    container_array<double>(g->values, box->values, g->counts, g->cap, 1, g->cells, o, d, g->d, edge, load);
    container_array<double>(g->influx, box->influx, g->counts, g->cap, 1, g->cells, o, d, g->d, edge, load);
    container_array<int32_t>(g->nbc, box->nb_counts, g->counts, g->cap, 1, g->cells, o, d, g->d, edge, load);
    container_array<int32_t>(g->nbids, box->nb_ids, g->counts, g->cap * g->maxnb, g->maxnb, g->cells, o, d, g->d, edge, load);
    if (load) g->dirty = true;
    return B200GEO_OK;
}

static bool container_box_ok(const b200geo_containergrid *g, const int32_t o[3], const int32_t d[3])
{
    for (int i = 0; i < 3; ++i)
        if (o[i] < 0 || d[i] < 1 || o[i] + d[i] > g->d[i]) return false;
    return true;
}

static bool container_box_complete(const b200geo_container_box *b)
{
    return b && b->counts && b->ids && b->values && b->influx && b->nb_counts && b->nb_ids;
}

int b200geo_containergrid_load(b200geo_containergrid *g, const int32_t o[3], const int32_t d[3], const b200geo_container_box *box, int, void *)
{
    if (!g || !o || !d) return fail(B200GEO_ERR_INVALID, "null argument");
    if (!container_box_complete(box)) return fail(B200GEO_ERR_INVALID, "a load needs every array of the box");
    if (!container_box_ok(g, o, d)) return fail(B200GEO_ERR_INVALID, "box outside the grid");
    return container_io(g, o, d, box, false, true);
}

int b200geo_containergrid_save(const b200geo_containergrid *g, const int32_t o[3], const int32_t d[3], const b200geo_container_box *box, int, void *)
{
    if (!g || !o || !d || !box) return fail(B200GEO_ERR_INVALID, "null argument");
    if (!container_box_ok(g, o, d)) return fail(B200GEO_ERR_INVALID, "box outside the grid");
    return container_io(const_cast<b200geo_containergrid *>(g), o, d, box, false, false);
}

int b200geo_containergrid_set_edge(b200geo_containergrid *g, const b200geo_container_box *cell)
{
    if (!g || !container_box_complete(cell)) return fail(B200GEO_ERR_INVALID, "the edge container needs every array");
    if (cell->counts[0] < 0 || cell->counts[0] > g->cap) return fail(B200GEO_ERR_OUT_OF_RANGE, "ContainerCell capacity exeeded");
    const int32_t o[3] = {0, 0, 0}, d[3] = {1, 1, 1};
    return container_io(g, o, d, cell, true, true);
}

int b200geo_containergrid_get_edge(const b200geo_containergrid *g, const b200geo_container_box *cell)
{
    if (!g || !cell) return fail(B200GEO_ERR_INVALID, "null argument");
    const int32_t o[3] = {0, 0, 0}, d[3] = {1, 1, 1};
    return container_io(const_cast<b200geo_containergrid *>(g), o, d, cell, true, false);
}

int b200geo_containergrid_step(b200geo_containergrid *g, uint32_t, uint32_t n_steps, void *)
{
    if (!g) return fail(B200GEO_ERR_INVALID, "null argument");
    if (n_steps == 0) return B200GEO_OK;
    uint64_t links = 0;
    for (size_t c = 0; c <= g->cells; ++c) {
        if (g->counts[c] < 0 || g->counts[c] > g->cap) return fail(B200GEO_ERR_OUT_OF_RANGE, "ContainerCell capacity exeeded");
        for (int s = 0; s < g->counts[c]; ++s) {
            if (s > 0 && g->ids[c * g->cap + s - 1] >= g->ids[c * g->cap + s])
                return fail(B200GEO_ERR_INVALID, "the ids of a container must ascend (ContainerCell::insert keeps them sorted)");
            if (c < g->cells) links += g->nbc[c * g->cap + s];
        }
    }
    size_t e = g->cells * g->cap;
    std::vector<double> out(e);
    int32_t missing = 0;
    int rc = oracle_container(g->desc.n_dims, g->desc.ghost_mode[0][0] == B200GEO_GHOST_WRAP, g->d[0], g->d[1], g->d[2], g->cap, g->maxnb,
                              (int)n_steps, g->counts.data(), g->ids.data(), g->values.data(), g->influx.data(), g->nbc.data(),
                              g->nbids.data(), &g->counts[g->cells], &g->ids[e], &g->values[e], out.data(), &missing);
    if (rc == -2) return fail(B200GEO_ERR_LOGIC, "id not found: could not find id " + std::to_string(missing) + " in neighborhood");
    if (rc) return fail(B200GEO_ERR_INVALID, "oracle_container failed");
    memcpy(g->values.data(), out.data(), e * sizeof(double));
    g->rebuilds += g->dirty;
    g->dirty = false;
    g->links = links;
    g->sweeps += n_steps;
    return B200GEO_OK;
}

int b200geo_containergrid_stats(const b200geo_containergrid *g, uint64_t out[6])
{
    if (!g || !out) return fail(B200GEO_ERR_INVALID, "null argument");
    uint64_t cargo = 0;
    for (size_t c = 0; c < g->cells; ++c) cargo += g->counts[c];
    out[0] = cargo;
    out[1] = g->links;
    out[2] = g->rebuilds;
    out[3] = g->sweeps;
    out[4] = 0;
    out[5] = 4 * g->links;
    return B200GEO_OK;
}

}
