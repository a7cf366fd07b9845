// libb200geo.so — grid life cycle, bulk I/O, step dispatch and halo plumbing behind the C ABI
// declared in include/b200geo.h. The hot kernels live in jacobi.cu, gol.cu, lbm.cu, region.cu.
#include "grid.h"

#include <atomic>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

namespace b200geo {

static const Tuning g_tuning_default = {0, -1, 0, 128, 0, 33, 0, 4, 0, 3, -1, 1, 0, 0, 0, 0, 0, 2, 14, 0, 0, 1, 0, 0};
Tuning g_tuning = g_tuning_default;
static thread_local std::string g_last_error;
static std::atomic<uint64_t> g_launches(0);

int fail(int status, const std::string& msg)
{
    g_last_error = msg;
    return status;
}

int check_cuda(cudaError_t e, const char *what)
{
    if (e == cudaSuccess) return 0;
    g_last_error = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what;
    return B200GEO_ERR_CUDA;
}

void count_launch(uint64_t n)
{
    g_launches += n;
}

static int64_t round_up(int64_t v, int64_t a)
{
    return (v + a - 1) / a * a;
}

static bool valid_box(const b200geo_grid *g, const int32_t o[3], const int32_t d[3])
{
    for (int i = 0; i < 3; ++i) {
        if (d[i] < 0 || o[i] < -g->g[i] || o[i] + d[i] > g->d[i] + g->g[i]) return false;
    }
    return true;
}

}

using namespace b200geo;

extern "C" {

const char *b200geo_version(void)
{
    return "b200geo 0.1 (sm_100a)";
}

const char *b200geo_last_error(void)
{
    return g_last_error.c_str();
}

int b200geo_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return check_cuda(e, "cudaGetDeviceCount");
    return n;
}

int b200geo_set_tuning(const char *key, int value)
{
    if (!key) return fail(B200GEO_ERR_INVALID, "null key");
    std::string k(key);
    if (k == "jacobi.zchunk") g_tuning.jacobi_zchunk = value < 0 ? g_tuning_default.jacobi_zchunk : value;
    else if (k == "jacobi.prefetch") g_tuning.jacobi_prefetch = value;
    else if (k == "gol.rows") g_tuning.gol_rows = value < 0 ? g_tuning_default.gol_rows : value;
    else if (k == "lbm.block") g_tuning.lbm_block = value < 0 ? g_tuning_default.lbm_block : value;
    else if (k == "jacobi.tb") g_tuning.jacobi_tb = value < 0 ? g_tuning_default.jacobi_tb : value;
    else if (k == "jacobi.tb_rows") g_tuning.jacobi_tb_rows = value < 0 ? g_tuning_default.jacobi_tb_rows : value;
    else if (k == "jacobi.tb_zchunk") g_tuning.jacobi_tb_zchunk = value < 0 ? g_tuning_default.jacobi_tb_zchunk : value;
    else if (k == "gol.bits") g_tuning.gol_bits = value < 0 ? g_tuning_default.gol_bits : value;
    else if (k == "gol.bits_rows") g_tuning.gol_bits_rows = value < 0 ? g_tuning_default.gol_bits_rows : value;
    else if (k == "nbody.kernel") g_tuning.nbody_kernel = value < 0 ? g_tuning_default.nbody_kernel : value;
    else if (k == "jacobi.pdl") g_tuning.jacobi_pdl = value;
    else if (k == "jacobi.tb_promo") g_tuning.jacobi_tb_promo = value < 0 ? g_tuning_default.jacobi_tb_promo : value;
    else if (k == "nbody.run") g_tuning.nbody_run = value < 0 ? g_tuning_default.nbody_run : value;
    else if (k == "nbody.threads") g_tuning.nbody_threads = value < 0 ? g_tuning_default.nbody_threads : value;
    else if (k == "jacobi.resident") g_tuning.jacobi_resident = value < 0 ? g_tuning_default.jacobi_resident : value;
    else if (k == "jacobi.tb_raster") g_tuning.jacobi_tb_raster = value < 0 ? g_tuning_default.jacobi_tb_raster : value;
    else if (k == "lbm.variant") g_tuning.lbm_variant = value < 0 ? g_tuning_default.lbm_variant : value;
    else if (k == "lbm.tb") g_tuning.lbm_tb = value < 0 ? g_tuning_default.lbm_tb : value;
    else if (k == "lbm.tb_rows") g_tuning.lbm_tb_rows = value < 0 ? g_tuning_default.lbm_tb_rows : value;
    else if (k == "lbm.tb_promo") g_tuning.lbm_tb_promo = value < 0 ? g_tuning_default.lbm_tb_promo : value;
    else if (k == "lbm.tb_warps") g_tuning.lbm_tb_warps = value < 0 ? g_tuning_default.lbm_tb_warps : value;
    else if (k == "lbm.tb_hints") g_tuning.lbm_tb_hints = value < 0 ? g_tuning_default.lbm_tb_hints : value;
    else if (k == "lbm.tb_zchunk") g_tuning.lbm_tb_zchunk = value < 0 ? g_tuning_default.lbm_tb_zchunk : value;
    else if (k == "container.kernel") g_tuning.container_kernel = value < 0 ? g_tuning_default.container_kernel : value;
    else return fail(B200GEO_ERR_INVALID, "unknown tuning key " + k);
    return B200GEO_OK;
}

uint64_t b200geo_launch_count(void)
{
    return g_launches.load();
}

// Geometry of the uniform element layout (b200geo_grid_create_uniform): ONE lead-in and ONE pitch, counted in
// elements, for all members, so that one element index addresses every member of a cell. The lead-in is 128 bytes
// of the narrowest member (rows of wider members then start on multiples of 128 bytes as well).
static void uniform_geometry(const b200geo_grid_desc *desc, int *lead, int64_t *pitch, int64_t *min_stride)
{
    int min_e = 8;
    for (int m = 0; m < desc->n_members; ++m) min_e = desc->member_bytes[m] < min_e ? desc->member_bytes[m] : min_e;
    if (min_e < 1) min_e = 1;
    *lead = 128 / min_e;
    *pitch = round_up((int64_t)*lead + desc->dim[0] + desc->ghost[0], *lead);
    int64_t elems = *pitch * (desc->dim[1] + 2 * desc->ghost[1]) * (desc->dim[2] + 2 * desc->ghost[2]);
    *min_stride = round_up(elems + *lead, 256);   // same slack behind the last row as the default layout
}

static int create_grid(const b200geo_grid_desc *desc, int device, int64_t uniform_stride, b200geo_grid **out);

int b200geo_grid_create(const b200geo_grid_desc *desc, int device, b200geo_grid **out)
{
    return create_grid(desc, device, 0, out);
}

int b200geo_grid_create_uniform(const b200geo_grid_desc *desc, int device, int64_t member_stride, b200geo_grid **out)
{
    if (member_stride <= 0 || member_stride % 256 != 0)
        return fail(B200GEO_ERR_INVALID, "member stride must be a positive multiple of 256 elements");
    return create_grid(desc, device, member_stride, out);
}

int b200geo_grid_uniform_min_stride(const b200geo_grid_desc *desc, int64_t *min_stride)
{
    if (!desc || !min_stride) return fail(B200GEO_ERR_INVALID, "null argument");
    if (desc->n_members < 1 || desc->n_members > B200GEO_MAX_MEMBERS)
        return fail(B200GEO_ERR_INVALID, "n_members out of range");
    for (int m = 0; m < desc->n_members; ++m) {
        int e = desc->member_bytes[m];
        if (e != 1 && e != 2 && e != 4 && e != 8)
            return fail(B200GEO_ERR_INVALID, "member size must be 1, 2, 4 or 8 bytes");
    }
    for (int i = 0; i < 3; ++i)
        if (desc->dim[i] < 1 || desc->ghost[i] < 0 || desc->ghost[i] > (i == 0 ? 16 : 65536))
            return fail(B200GEO_ERR_INVALID, "grid dimension or ghost width out of range");
    int lead;
    int64_t pitch;
    uniform_geometry(desc, &lead, &pitch, min_stride);
    return B200GEO_OK;
}

// Checks a grid description and lays out its member arrays (default layout: uniform_stride = 0). No device needed.
static int plan_layout(const b200geo_grid_desc *desc, int64_t uniform_stride, MemberLayout *layout, int64_t *buffer_bytes,
                       int *cell_bytes)
{
    if (!desc) return fail(B200GEO_ERR_INVALID, "null argument");
    if (desc->n_members < 1 || desc->n_members > B200GEO_MAX_MEMBERS)
        return fail(B200GEO_ERR_INVALID, "n_members out of range");
    for (int i = 0; i < 3; ++i) {
        if (desc->dim[i] < 1) return fail(B200GEO_ERR_INVALID, "grid dimension must be >= 1");
        // x ghosts live in the 128-byte lead-in of a row; along y and z a ghost zone may be as wide as a whole run is
        // long (streamed runs on slabs: no exchange while the wavefront passes)
        if (desc->ghost[i] < 0 || desc->ghost[i] > (i == 0 ? 16 : 65536)) return fail(B200GEO_ERR_INVALID, "ghost width out of range");
        for (int s = 0; s < 2; ++s) {
            int mode = desc->ghost_mode[i][s];
            if (mode < B200GEO_GHOST_EDGE || mode > B200GEO_GHOST_PEER)
                return fail(B200GEO_ERR_INVALID, "bad ghost mode");
            int slab = (desc->dim[2] == 1 && desc->ghost[2] == 0) ? 1 : 2;
            if (mode == B200GEO_GHOST_PEER && i != slab)
                return fail(B200GEO_ERR_LOGIC, "PEER ghost layers are supported on the last axis only (slab partition)");
            if (mode == B200GEO_GHOST_WRAP && desc->ghost[i] > desc->dim[i])
                return fail(B200GEO_ERR_INVALID, "wrap ghost wider than the grid");
        }
        if ((desc->ghost_mode[i][0] == B200GEO_GHOST_WRAP) != (desc->ghost_mode[i][1] == B200GEO_GHOST_WRAP))
            return fail(B200GEO_ERR_INVALID, "WRAP must be set on both sides of an axis");
    }
    for (int m = 0; m < desc->n_members; ++m) {
        int e = desc->member_bytes[m];
        if (e != 1 && e != 2 && e != 4 && e != 8)
            return fail(B200GEO_ERR_INVALID, "member size must be 1, 2, 4 or 8 bytes");
        if (desc->ghost[0] > 128 / e) return fail(B200GEO_ERR_INVALID, "x ghost wider than the 128-byte lead-in");
    }
    int ulead = 0;
    int64_t upitch = 0, umin = 0;
    if (uniform_stride > 0) {
        uniform_geometry(desc, &ulead, &upitch, &umin);
        if (uniform_stride < umin) return fail(B200GEO_ERR_INVALID, "member stride smaller than the padded grid");
    }
    const int *d = desc->dim, *gh = desc->ghost;
    int64_t off = 0;
    int cell = 0;
    for (int m = 0; m < desc->n_members; ++m) {
        int e = desc->member_bytes[m];
        MemberLayout& L = layout[m];
        L.elem = e;
        L.lead = uniform_stride > 0 ? ulead : 128 / e;
        if (gh[0] > L.lead) return fail(B200GEO_ERR_INVALID, "x ghost wider than the 128-byte lead-in");
        L.pitch = uniform_stride > 0 ? upitch : round_up((int64_t)(L.lead + d[0] + gh[0]) * e, 128) / e;
        L.plane = L.pitch * (d[1] + 2 * gh[1]);
        L.origin = (int64_t)gh[2] * L.plane + (int64_t)gh[1] * L.pitch + L.lead;
        // 128 B of slack: vector accesses of partially valid groups may touch the bytes just past
        // the last row (never stored to, values never used)
        L.bytes = round_up(L.plane * (d[2] + 2 * gh[2]) * e + 128, 256);
        // uniform layout: every member array has member_stride elements, so member m starts member_stride x
        // (bytes of the members before it) into the buffer — LibFlatArray's DIM_PROD x offset<CELL, m>
        if (uniform_stride > 0) L.bytes = uniform_stride * e;
        L.offset = off;
        L.edge_offset = cell;
        off += L.bytes;
        cell += e;
    }
    *buffer_bytes = off;
    *cell_bytes = cell;
    return B200GEO_OK;
}

int b200geo_grid_plan(const b200geo_grid_desc *desc, int64_t member_stride, int64_t *layout, int64_t *buffer_bytes)
{
    if (!layout || !buffer_bytes) return fail(B200GEO_ERR_INVALID, "null argument");
    if (member_stride < 0 || member_stride % 256 != 0)
        return fail(B200GEO_ERR_INVALID, "member stride must be a multiple of 256 elements (0 = default layout)");
    MemberLayout m[B200GEO_MAX_MEMBERS];
    int cell = 0;
    int rc = plan_layout(desc, member_stride, m, buffer_bytes, &cell);
    if (rc) return rc;
    for (int i = 0; i < desc->n_members; ++i) {
        int64_t *row = layout + 7 * i;
        row[0] = m[i].elem; row[1] = m[i].lead; row[2] = m[i].pitch; row[3] = m[i].plane;
        row[4] = m[i].origin; row[5] = m[i].bytes; row[6] = m[i].offset;
    }
    return B200GEO_OK;
}

static int create_grid(const b200geo_grid_desc *desc, int device, int64_t uniform_stride, b200geo_grid **out)
{
    if (!out) return fail(B200GEO_ERR_INVALID, "null argument");
    MemberLayout layout[B200GEO_MAX_MEMBERS];
    int64_t off = 0;
    int cell = 0;
    int rc = plan_layout(desc, uniform_stride, layout, &off, &cell);
    if (rc) return rc;
    B200GEO_CUDA(cudaSetDevice(device));

    b200geo_grid *g = new (std::nothrow) b200geo_grid();
    if (!g) return fail(B200GEO_ERR_NOMEM, "out of host memory");
    memset(g, 0, sizeof(*g));
    g->desc = *desc;
    g->device = device;
    g->n = desc->n_members;
    for (int i = 0; i < 3; ++i) {
        g->d[i] = desc->dim[i];
        g->g[i] = desc->ghost[i];
    }
    g->slab_axis = (g->d[2] == 1 && g->g[2] == 0) ? 1 : 2;
    g->uniform_stride = uniform_stride;
    for (int m = 0; m < g->n; ++m) g->m[m] = layout[m];
    g->buffer_bytes = off;
    g->cell_bytes = cell;
    for (int b = 0; b < 2; ++b) {
        cudaError_t e = cudaMalloc((void **)&g->buf[b], (size_t)off);
        if (e != cudaSuccess) {
            if (b == 1) cudaFree(g->buf[0]);
            delete g;
            cudaGetLastError();
            return fail(B200GEO_ERR_NOMEM, std::string("cudaMalloc failed: ") + cudaGetErrorString(e));
        }
    }
    // zero-fill: the edge cell defaults to all-zero bytes until set_edge is called
    cudaMemset(g->buf[0], 0, (size_t)off);
    cudaMemset(g->buf[1], 0, (size_t)off);
    for (int i = 0; i < 4; ++i) cudaEventCreate(&g->ev[i]);
    *out = g;
    return B200GEO_OK;
}

int b200geo_grid_member_stride(const b200geo_grid *g, int64_t *member_stride)
{
    if (!g || !member_stride) return fail(B200GEO_ERR_INVALID, "null argument");
    *member_stride = g->uniform_stride;
    return B200GEO_OK;
}

int b200geo_grid_destroy(b200geo_grid *g)
{
    if (!g) return B200GEO_OK;
    cudaSetDevice(g->device);
    for (int s = 0; s < 2; ++s)
        for (int w = 0; w < 2; ++w)
            if (g->peer_buf[s][w]) cudaIpcCloseMemHandle(g->peer_buf[s][w]);
    cudaFree(g->buf[0]);
    cudaFree(g->buf[1]);
    if (g->scratch) cudaFree(g->scratch);
    if (g->io_staging) cudaFree(g->io_staging);
    if (g->bits) cudaFree(g->bits);
    for (int i = 0; i < 4; ++i) cudaEventDestroy(g->ev[i]);
    delete g;
    return B200GEO_OK;
}

int b200geo_grid_buffer_bytes(const b200geo_grid *g, uint64_t *bytes)
{
    if (!g || !bytes) return fail(B200GEO_ERR_INVALID, "null argument");
    *bytes = (uint64_t)g->buffer_bytes;
    return B200GEO_OK;
}

int b200geo_grid_device(const b200geo_grid *g, int *device)
{
    if (!g || !device) return fail(B200GEO_ERR_INVALID, "null argument");
    *device = g->device;
    return B200GEO_OK;
}

int b200geo_grid_layout(const b200geo_grid *g, int member, int64_t *pitch_x, int64_t *pitch_plane,
                        int64_t *origin_offset)
{
    if (!g || member < 0 || member >= g->n) return fail(B200GEO_ERR_INVALID, "bad member");
    if (pitch_x) *pitch_x = g->m[member].pitch;
    if (pitch_plane) *pitch_plane = g->m[member].plane;
    if (origin_offset) *origin_offset = g->m[member].origin;
    return B200GEO_OK;
}

int b200geo_grid_member_ptr(const b200geo_grid *g, int member, int which, void **ptr)
{
    if (!g || !ptr || member < 0 || member >= g->n || (which != 0 && which != 1))
        return fail(B200GEO_ERR_INVALID, "bad member");
    *ptr = g->member_ptr(member, which);
    return B200GEO_OK;
}

int b200geo_grid_set_edge(b200geo_grid *g, const void *cell, void *stream)
{
    if (!g || !cell) return fail(B200GEO_ERR_INVALID, "null argument");
    memcpy(g->edge, cell, g->cell_bytes);
    B200GEO_CUDA(cudaSetDevice(g->device));
    int rc = fill_edge(g, 0, (cudaStream_t)stream);
    if (rc) return rc;
    return fill_edge(g, 1, (cudaStream_t)stream);
}

int b200geo_grid_get_edge(const b200geo_grid *g, void *cell)
{
    if (!g || !cell) return fail(B200GEO_ERR_INVALID, "null argument");
    memcpy(cell, g->edge, g->cell_bytes);
    return B200GEO_OK;
}

static int member_copy(const b200geo_grid *g, int member, const int32_t o[3], const int32_t d[3],
                       void *dense, int location, bool to_grid, int which, cudaStream_t s)
{
    const MemberLayout& L = g->m[member];
    if (d[0] == 0 || d[1] == 0 || d[2] == 0) return B200GEO_OK;
    cudaMemcpy3DParms p;
    memset(&p, 0, sizeof(p));
    cudaPitchedPtr dense_ptr = make_cudaPitchedPtr(dense, (size_t)d[0] * L.elem, (size_t)d[0] * L.elem, d[1]);
    cudaPitchedPtr grid_ptr = make_cudaPitchedPtr(g->member_ptr(member, which), (size_t)L.pitch * L.elem,
                                                  (size_t)L.pitch * L.elem, g->d[1] + 2 * g->g[1]);
    cudaPos grid_pos = make_cudaPos((size_t)(L.lead + o[0]) * L.elem, o[1] + g->g[1], o[2] + g->g[2]);
    if (to_grid) {
        p.srcPtr = dense_ptr;
        p.dstPtr = grid_ptr;
        p.dstPos = grid_pos;
        p.kind = location == B200GEO_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    } else {
        p.srcPtr = grid_ptr;
        p.srcPos = grid_pos;
        p.dstPtr = dense_ptr;
        p.kind = location == B200GEO_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    }
    p.extent = make_cudaExtent((size_t)d[0] * L.elem, d[1], d[2]);
    B200GEO_CUDA(cudaMemcpy3DAsync(&p, s));
    return B200GEO_OK;
}

int b200geo_grid_load_member(b200geo_grid *g, int member, const int32_t origin[3], const int32_t dim[3],
                             const void *src, int location, int both, void *stream)
{
    if (!g || !src || !origin || !dim) return fail(B200GEO_ERR_INVALID, "null argument");
    if (member < 0 || member >= g->n) return fail(B200GEO_ERR_INVALID, "bad member index");
    if (!valid_box(g, origin, dim)) return fail(B200GEO_ERR_INVALID, "box outside the grid");
    B200GEO_CUDA(cudaSetDevice(g->device));
    int rc = member_copy(g, member, origin, dim, const_cast<void *>(src), location, true, 0, (cudaStream_t)stream);
    if (rc == 0 && both) {
        // second buffer: device-to-device copy of the same box, not a second trip over PCIe
        const MemberLayout& L = g->m[member];
        cudaMemcpy3DParms p;
        memset(&p, 0, sizeof(p));
        size_t pitch = (size_t)L.pitch * L.elem;
        p.srcPtr = make_cudaPitchedPtr(g->member_ptr(member, 0), pitch, pitch, g->d[1] + 2 * g->g[1]);
        p.dstPtr = make_cudaPitchedPtr(g->member_ptr(member, 1), pitch, pitch, g->d[1] + 2 * g->g[1]);
        p.srcPos = p.dstPos = make_cudaPos((size_t)(L.lead + origin[0]) * L.elem, origin[1] + g->g[1], origin[2] + g->g[2]);
        p.extent = make_cudaExtent((size_t)dim[0] * L.elem, dim[1], dim[2]);
        p.kind = cudaMemcpyDeviceToDevice;
        if (dim[0] > 0 && dim[1] > 0 && dim[2] > 0) rc = check_cuda(cudaMemcpy3DAsync(&p, (cudaStream_t)stream), "cudaMemcpy3DAsync");
    }
    return rc;
}

int b200geo_grid_save_member(const b200geo_grid *g, int member, const int32_t origin[3], const int32_t dim[3],
                             void *dst, int location, void *stream)
{
    if (!g || !dst || !origin || !dim) return fail(B200GEO_ERR_INVALID, "null argument");
    if (member < 0 || member >= g->n) return fail(B200GEO_ERR_INVALID, "bad member index");
    if (!valid_box(g, origin, dim)) return fail(B200GEO_ERR_INVALID, "box outside the grid");
    B200GEO_CUDA(cudaSetDevice(g->device));
    return member_copy(g, member, origin, dim, dst, location, false, 0, (cudaStream_t)stream);
}

static int region_io(b200geo_grid *g, const int32_t *streaks, int n_streaks, void *buf, int location,
                     bool save, int both, cudaStream_t s)
{
    if (!g || (n_streaks > 0 && (!streaks || !buf))) return fail(B200GEO_ERR_INVALID, "null argument");
    if (n_streaks <= 0) return B200GEO_OK;
    int64_t count = 0;
    for (int i = 0; i < n_streaks; ++i) {
        const int32_t *k = streaks + 4 * i;
        if (k[3] < k[0] || k[0] < -g->g[0] || k[3] > g->d[0] + g->g[0] ||
            k[1] < -g->g[1] || k[1] >= g->d[1] + g->g[1] || k[2] < -g->g[2] || k[2] >= g->d[2] + g->g[2])
            return fail(B200GEO_ERR_INVALID, "streak outside the grid");
        count += k[3] - k[0];
    }
    B200GEO_CUDA(cudaSetDevice(g->device));
    char *dev = (char *)buf;
    size_t bytes = (size_t)count * g->cell_bytes;
    const bool staged = location == B200GEO_HOST;
    if (staged) {
        // Initializers and Writers that go cell by cell (GridBase::set / get in a loop) come through here once
        // per call: the staging buffer is kept, so such a call is two small copies and one launch
        if (g->io_staging_bytes < bytes) {
            if (g->io_staging) {
                cudaStreamSynchronize(s);
                cudaFree(g->io_staging);
                g->io_staging = 0;
                g->io_staging_bytes = 0;
            }
            size_t want = bytes < 4096 ? 4096 : bytes;
            cudaError_t e = cudaMalloc((void **)&g->io_staging, want);
            if (e != cudaSuccess) {
                cudaGetLastError();
                return fail(B200GEO_ERR_NOMEM, std::string("cudaMalloc failed: ") + cudaGetErrorString(e));
            }
            g->io_staging_bytes = want;
        }
        dev = g->io_staging;
        // pageable host memory: the copy has left the caller's buffer when cudaMemcpyAsync returns
        if (!save) B200GEO_CUDA(cudaMemcpyAsync(dev, buf, bytes, cudaMemcpyHostToDevice, s));
    }
    int rc = copy_region(g, streaks, n_streaks, dev, count, save, 0, s);
    if (rc == 0 && !save && both) rc = copy_region(g, streaks, n_streaks, dev, count, false, 1, s);
    if (rc == 0 && staged && save) {
        // the caller reads its buffer right after the call (and the next call reuses the staging buffer)
        rc = check_cuda(cudaMemcpyAsync(buf, dev, bytes, cudaMemcpyDeviceToHost, s), "cudaMemcpyAsync");
        if (rc == 0) rc = check_cuda(cudaStreamSynchronize(s), "cudaStreamSynchronize");
    }
    return rc;
}

int b200geo_grid_load_region(b200geo_grid *g, const int32_t *streaks, int n_streaks,
                             const void *buf, int location, int both, void *stream)
{
    return region_io(g, streaks, n_streaks, const_cast<void *>(buf), location, false, both, (cudaStream_t)stream);
}

int b200geo_grid_save_region(const b200geo_grid *g, const int32_t *streaks, int n_streaks,
                             void *buf, int location, void *stream)
{
    return region_io(const_cast<b200geo_grid *>(g), streaks, n_streaks, buf, location, true, 0, (cudaStream_t)stream);
}

// LBM params: int32 macroscopic store mode: 0 (default) = only on the last sweep of this call, 1 = on every sweep
// (what the reference cell does; liquid cells overwrite them sweep after sweep, wall cells never write them: the grid
// after the call is the same as with mode 0), 2 = never (caller is mid-run)
static bool lbm_stores_macroscopic(const void *params, bool last)
{
    int mode = params ? *(const int32_t *)params : 0;
    return mode == 1 || (mode == 0 && last);
}

static int dispatch(b200geo_grid *g, int kernel, const void *params, const Box& box, bool last, cudaStream_t s)
{
    if (box.x1 <= box.x0 || box.y1 <= box.y0 || box.z1 <= box.z0) return B200GEO_OK;
    switch (kernel) {
    case B200GEO_KERNEL_JACOBI6:
        return sweep_jacobi(g, 6, box, s);
    case B200GEO_KERNEL_JACOBI7:
        return sweep_jacobi(g, 7, box, s);
    case B200GEO_KERNEL_JACOBI27:
        return sweep_jacobi(g, 27, box, s);
    case B200GEO_KERNEL_GOL:
        return sweep_gol(g, box, s);
    case B200GEO_KERNEL_LBM_D3Q19:
        return sweep_lbm(g, box, lbm_stores_macroscopic(params, last), s);
    default:
        return fail(B200GEO_ERR_LOGIC, "no kernel bound for this id");
    }
}

static int check_kernel_grid(const b200geo_grid *g, int kernel)
{
    switch (kernel) {
    case B200GEO_KERNEL_JACOBI6:
    case B200GEO_KERNEL_JACOBI7:
    case B200GEO_KERNEL_JACOBI27:
        if (g->n != 1 || g->m[0].elem != 8) return fail(B200GEO_ERR_INVALID, "Jacobi kernels need one f64 member");
        for (int i = 0; i < 3; ++i)
            if (g->g[i] < 1) return fail(B200GEO_ERR_INVALID, "Jacobi kernels need ghost width >= 1 on all axes");
        return 0;
    case B200GEO_KERNEL_GOL:
        if (g->n != 1 || g->m[0].elem != 1) return fail(B200GEO_ERR_INVALID, "GoL kernel needs one 1-byte member");
        if (g->d[2] != 1 || g->g[0] < 1 || g->g[1] < 1) return fail(B200GEO_ERR_INVALID, "GoL kernel needs a 2-D grid with ghost width >= 1");
        return 0;
    case B200GEO_KERNEL_LBM_D3Q19:
        if (g->n != 24) return fail(B200GEO_ERR_INVALID, "LBM kernel needs 24 members");
        for (int m = 0; m < 24; ++m)
            if (g->m[m].elem != 4) return fail(B200GEO_ERR_INVALID, "LBM kernel needs 4-byte members");
        for (int i = 0; i < 3; ++i)
            if (g->g[i] < 1) return fail(B200GEO_ERR_INVALID, "LBM kernel needs ghost width >= 1 on all axes");
        return 0;
    default:
        return fail(B200GEO_ERR_LOGIC, "no kernel bound for this id");
    }
}

int b200geo_step(b200geo_grid *g, int kernel, const void *params, uint32_t first_nano_step,
                 uint32_t n_steps, void *stream)
{
    (void)first_nano_step;  // none of the bound models depends on the nano step index
    if (!g) return fail(B200GEO_ERR_INVALID, "null grid");
    int rc = check_kernel_grid(g, kernel);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    B200GEO_CUDA(cudaSetDevice(g->device));
    if (g->stats_on) cudaEventRecord(g->ev[0], s);
    const int a = g->slab_axis;
    // Game of Life, several sweeps in one call: pack to one bit per cell, sweep the packed copy (it
    // stays in L2), unpack. The byte grid, its ghost ring and `cur` are as after n_steps byte sweeps.
    if (kernel == B200GEO_KERNEL_GOL && g_tuning.gol_bits > 0 && n_steps >= (uint32_t)g_tuning.gol_bits && gol_bits_applicable(g)) {
        rc = sweep_gol_bits(g, n_steps, s);
        if (rc) return rc;
        g->sweeps += n_steps;
        n_steps = 0;
    }
    const bool jacobi = kernel == B200GEO_KERNEL_JACOBI6 || kernel == B200GEO_KERNEL_JACOBI7 || kernel == B200GEO_KERNEL_JACOBI27;
    const bool lbm = kernel == B200GEO_KERNEL_LBM_D3Q19;
    // Small grids (a few million cells: BASELINE.json configs[0]): all sweeps of this call in ONE cooperative launch
    // that keeps every brick of the grid resident in an SM (jacobi_resident.cu). Off unless "jacobi.resident" = 1:
    // measured no faster than the streaming kernel with dependent launches (profiles/r3i_r3j_r3k)
    if (jacobi && n_steps >= 2 && g_tuning.jacobi_resident > 0) {
        const int kind = kernel == B200GEO_KERNEL_JACOBI6 ? 6 : kernel == B200GEO_KERNEL_JACOBI7 ? 7 : 27;
        const int planes = jacobi_resident_planes(g, kind);
        if (planes > 0) {
            rc = sweep_jacobi_resident(g, kind, planes, (int)n_steps, s);
            if (rc) return rc;
            g->cur ^= (int)(n_steps & 1);
            g->sweeps += n_steps;
            n_steps = 0;
        }
    }
    for (uint32_t t = 0; t < n_steps;) {
        // sweeps fused into this launch: the temporal-blocked Jacobi kernel takes `depth` sweeps per
        // HBM round trip when enough valid ghost cells are there (WRAP: ghost width, PEER: what the
        // last exchange delivered); everything else is one sweep per launch
        int depth = 1;
        // 0 = automatic (measured on B200, profiles/r1s_tuning.md): the 27-point kernel is fastest with two
        // fused sweeps (deeper blocking runs out of registers), the 6/7-point kernels with four
        // the LBM kernel fuses two (lbm_tb.cu)
        const int tb = lbm ? (g_tuning.lbm_tb >= 2 ? 2 : 1) :
            g_tuning.jacobi_tb != 0 ? g_tuning.jacobi_tb : (kernel == B200GEO_KERNEL_JACOBI27 ? 2 : 4);
        if ((jacobi || lbm) && tb > 1) {
            depth = tb > 4 ? 4 : tb;
            if ((uint32_t)depth > n_steps - t) depth = (int)(n_steps - t);
            for (int i = 0; i < 3; ++i)
                for (int side = 0; side < 2; ++side) {
                    int mode = g->desc.ghost_mode[i][side];
                    if (mode == B200GEO_GHOST_WRAP && g->g[i] < depth) depth = g->g[i];
                    if (mode == B200GEO_GHOST_PEER && g->peer_valid[side] < depth) depth = g->peer_valid[side];
                }
            if (depth < 1) depth = 1;
        }
        Box box = {0, 0, 0, g->d[0], g->d[1], g->d[2]};
        for (int side = 0; side < 2; ++side) {
            if (g->desc.ghost_mode[a][side] != B200GEO_GHOST_PEER) continue;
            if (g->peer_valid[side] < 1)
                return fail(B200GEO_ERR_LOGIC, "ghost zone exhausted: exchange halos before stepping");
            int extra = g->peer_valid[side] - depth;
            if (a == 2) { if (side == 0) box.z0 -= extra; else box.z1 += extra; }
            else        { if (side == 0) box.y0 -= extra; else box.y1 += extra; }
        }
        rc = refresh_wrap(g, s);
        if (rc) return rc;
        if (depth > 1 && lbm) rc = sweep_lbm_tb2(g, box, lbm_stores_macroscopic(params, t + depth == n_steps), s);
        else if (depth > 1) rc = sweep_jacobi_tb(g, kernel == B200GEO_KERNEL_JACOBI6 ? 6 : kernel == B200GEO_KERNEL_JACOBI7 ? 7 : 27, depth, box, s);
        else rc = dispatch(g, kernel, params, box, t + 1 == n_steps, s);
        if (rc) return rc;
        g->cur ^= 1;
        for (int side = 0; side < 2; ++side)
            if (g->desc.ghost_mode[a][side] == B200GEO_GHOST_PEER) g->peer_valid[side] -= depth;
        g->sweeps += depth;
        t += depth;
    }
    if (g->stats_on) {
        cudaEventRecord(g->ev[1], s);
        B200GEO_CUDA(cudaEventSynchronize(g->ev[1]));
        float ms = 0;
        cudaEventElapsedTime(&ms, g->ev[0], g->ev[1]);
        g->t_update += 1e-3 * ms;
    }
    return B200GEO_OK;
}

int b200geo_update_box(b200geo_grid *g, int kernel, const void *params, uint32_t nano_step,
                       const int32_t origin[3], const int32_t dim[3], void *stream)
{
    (void)nano_step;
    if (!g || !origin || !dim) return fail(B200GEO_ERR_INVALID, "null argument");
    int rc = check_kernel_grid(g, kernel);
    if (rc) return rc;
    // the updated box may reach into PEER ghost planes but must leave one ring of readable cells
    for (int i = 0; i < 3; ++i) {
        bool slab = i == g->slab_axis && g->g[i] > 0;
        int lo = slab ? -(g->g[i] - 1) : 0, hi = g->d[i] + (slab ? g->g[i] - 1 : 0);
        if (dim[i] < 0 || origin[i] < lo || origin[i] + dim[i] > hi) return fail(B200GEO_ERR_INVALID, "box outside the updatable area");
    }
    B200GEO_CUDA(cudaSetDevice(g->device));
    Box box = {origin[0], origin[1], origin[2], origin[0] + dim[0], origin[1] + dim[1], origin[2] + dim[2]};
    return dispatch(g, kernel, params, box, true, (cudaStream_t)stream);
}

int b200geo_update_box_n(b200geo_grid *g, int kernel, const void *params, uint32_t nano_step,
                         const int32_t origin[3], const int32_t dim[3], uint32_t n_sweeps, void *stream)
{
    if (n_sweeps == 1) return b200geo_update_box(g, kernel, params, nano_step, origin, dim, stream);
    if (!g || !origin || !dim) return fail(B200GEO_ERR_INVALID, "null argument");
    int rc = check_kernel_grid(g, kernel);
    if (rc) return rc;
    const bool lbm = kernel == B200GEO_KERNEL_LBM_D3Q19;
    if (!lbm && kernel != B200GEO_KERNEL_JACOBI6 && kernel != B200GEO_KERNEL_JACOBI7 && kernel != B200GEO_KERNEL_JACOBI27)
        return fail(B200GEO_ERR_LOGIC, "this kernel family cannot fuse sweeps");
    if (n_sweeps < 1 || n_sweeps > (lbm ? 2u : 4u)) return fail(B200GEO_ERR_INVALID, lbm ? "n_sweeps must be 1..2" : "n_sweeps must be 1..4");
    const int n = (int)n_sweeps;
    for (int i = 0; i < 3; ++i) {
        bool slab = i == g->slab_axis && g->g[i] > 0;
        int lo = slab ? -(g->g[i] - n) : 0, hi = g->d[i] + (slab ? g->g[i] - n : 0);
        if (lo > 0) lo = 0;
        if (hi < g->d[i]) hi = g->d[i];
        if (dim[i] < 0 || origin[i] < lo || origin[i] + dim[i] > hi) return fail(B200GEO_ERR_INVALID, "box outside the updatable area");
        for (int side = 0; side < 2; ++side)
            if (g->desc.ghost_mode[i][side] != B200GEO_GHOST_EDGE && g->g[i] < n)
                return fail(B200GEO_ERR_INVALID, "ghost zone narrower than n_sweeps");
    }
    if (dim[0] == 0 || dim[1] == 0 || dim[2] == 0) return B200GEO_OK;
    B200GEO_CUDA(cudaSetDevice(g->device));
    Box box = {origin[0], origin[1], origin[2], origin[0] + dim[0], origin[1] + dim[1], origin[2] + dim[2]};
    if (lbm) return sweep_lbm_tb2(g, box, lbm_stores_macroscopic(params, true), (cudaStream_t)stream);
    return sweep_jacobi_tb(g, kernel == B200GEO_KERNEL_JACOBI6 ? 6 : kernel == B200GEO_KERNEL_JACOBI7 ? 7 : 27, n, box,
                           (cudaStream_t)stream);
}

int b200geo_swap(b200geo_grid *g)
{
    if (!g) return fail(B200GEO_ERR_INVALID, "null grid");
    g->cur ^= 1;
    ++g->sweeps;
    return B200GEO_OK;
}

int b200geo_refresh_ghosts(b200geo_grid *g, void *stream)
{
    if (!g) return fail(B200GEO_ERR_INVALID, "null grid");
    B200GEO_CUDA(cudaSetDevice(g->device));
    return refresh_wrap(g, (cudaStream_t)stream);
}

int b200geo_sync(void *stream)
{
    B200GEO_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return B200GEO_OK;
}

int b200geo_stream_create(int device, void **stream)
{
    if (!stream) return fail(B200GEO_ERR_INVALID, "null argument");
    B200GEO_CUDA(cudaSetDevice(device));
    cudaStream_t s;
    B200GEO_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *stream = (void *)s;
    return B200GEO_OK;
}

int b200geo_stream_destroy(int device, void *stream)
{
    B200GEO_CUDA(cudaSetDevice(device));
    B200GEO_CUDA(cudaStreamDestroy((cudaStream_t)stream));
    return B200GEO_OK;
}

int b200geo_stream_wait(int device, void *waiter, void *signaller)
{
    B200GEO_CUDA(cudaSetDevice(device));
    cudaEvent_t ev;
    B200GEO_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    int rc = check_cuda(cudaEventRecord(ev, (cudaStream_t)signaller), "cudaEventRecord");
    if (rc == 0) rc = check_cuda(cudaStreamWaitEvent((cudaStream_t)waiter, ev, 0), "cudaStreamWaitEvent");
    cudaEventDestroy(ev);  // released by the runtime once the recorded work has completed
    return rc;
}

int b200geo_grid_sync(const b200geo_grid *g, void *stream)
{
    if (!g) return fail(B200GEO_ERR_INVALID, "null grid");
    B200GEO_CUDA(cudaSetDevice(g->device));
    B200GEO_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return B200GEO_OK;
}

int b200geo_device_alloc(int device, uint64_t bytes, void **ptr)
{
    if (!ptr) return fail(B200GEO_ERR_INVALID, "null argument");
    *ptr = 0;
    B200GEO_CUDA(cudaSetDevice(device));
    cudaError_t e = cudaMalloc(ptr, bytes ? (size_t)bytes : 1);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(B200GEO_ERR_NOMEM, std::string("cudaMalloc failed: ") + cudaGetErrorString(e));
    }
    return B200GEO_OK;
}

int b200geo_device_free(int device, void *ptr)
{
    if (!ptr) return B200GEO_OK;
    B200GEO_CUDA(cudaSetDevice(device));
    B200GEO_CUDA(cudaFree(ptr));
    return B200GEO_OK;
}

int b200geo_device_copy(int dst_device, void *dst, int src_device, const void *src, uint64_t bytes)
{
    if (bytes == 0) return B200GEO_OK;
    if (!dst || !src) return fail(B200GEO_ERR_INVALID, "null argument");
    if (dst_device >= 0 && src_device >= 0 && dst_device != src_device) {
        B200GEO_CUDA(cudaSetDevice(dst_device));
        B200GEO_CUDA(cudaMemcpyPeer(dst, dst_device, src, src_device, (size_t)bytes));
        return B200GEO_OK;
    }
    B200GEO_CUDA(cudaSetDevice(dst_device >= 0 ? dst_device : src_device >= 0 ? src_device : 0));
    const cudaMemcpyKind kind = dst_device >= 0 ? (src_device >= 0 ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice) :
                                                  (src_device >= 0 ? cudaMemcpyDeviceToHost : cudaMemcpyHostToHost);
    B200GEO_CUDA(cudaMemcpy(dst, src, (size_t)bytes, kind));
    return B200GEO_OK;
}

int b200geo_host_alloc(uint64_t bytes, void **ptr)
{
    if (!ptr) return fail(B200GEO_ERR_INVALID, "null argument");
    *ptr = 0;
    cudaError_t e = cudaMallocHost(ptr, bytes ? (size_t)bytes : 1);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(B200GEO_ERR_NOMEM, std::string("cudaMallocHost failed: ") + cudaGetErrorString(e));
    }
    return B200GEO_OK;
}

int b200geo_host_free(void *ptr)
{
    if (!ptr) return B200GEO_OK;
    B200GEO_CUDA(cudaFreeHost(ptr));
    return B200GEO_OK;
}

int b200geo_halo_block(const b200geo_grid *g, int member, int side, int kind, int width,
                       void **ptr, uint64_t *bytes)
{
    return b200geo_halo_block_in(g, member, side, kind, width, 0, ptr, bytes);
}

int b200geo_halo_block_in(const b200geo_grid *g, int member, int side, int kind, int width, int which,
                          void **ptr, uint64_t *bytes)
{
    if (!g || !ptr || !bytes) return fail(B200GEO_ERR_INVALID, "null argument");
    if (member < 0 || member >= g->n || (side != 0 && side != 1) || (kind != 0 && kind != 1) || (which != 0 && which != 1))
        return fail(B200GEO_ERR_INVALID, "bad member/side/kind");
    const int a = g->slab_axis, gz = g->g[a], nz = g->d[a];
    if (width < 1 || width > gz || width > nz) return fail(B200GEO_ERR_INVALID, "bad halo width");
    const MemberLayout& L = g->m[member];
    int64_t zplane;  // padded slice index of the first slice of the block
    if (kind == 0) zplane = side == 0 ? gz : gz + nz - width;
    else           zplane = side == 0 ? gz - width : gz + nz;
    *ptr = g->member_ptr(member, which) + zplane * g->slice_elems(member) * L.elem;
    *bytes = (uint64_t)width * g->slice_elems(member) * L.elem;
    return B200GEO_OK;
}

int b200geo_halo_mark_valid(b200geo_grid *g, int side, int width)
{
    if (!g || (side != 0 && side != 1) || width < 0 || width > g->g[g->slab_axis]) return fail(B200GEO_ERR_INVALID, "bad side/width");
    g->peer_valid[side] = width;
    return B200GEO_OK;
}

int b200geo_grid_ipc_export(const b200geo_grid *g, int which, void *handle64)
{
    if (!g || !handle64 || (which != 0 && which != 1)) return fail(B200GEO_ERR_INVALID, "bad argument");
    B200GEO_CUDA(cudaSetDevice(g->device));
    cudaIpcMemHandle_t h;
    // `which` is absolute here (buffer 0 / 1), not relative to the current buffer
    B200GEO_CUDA(cudaIpcGetMemHandle(&h, g->buf[which]));
    memcpy(handle64, &h, sizeof(h));
    return B200GEO_OK;
}

int b200geo_grid_ipc_open(b200geo_grid *g, int side, int which, const void *handle64)
{
    if (!g || !handle64 || (side != 0 && side != 1) || (which != 0 && which != 1))
        return fail(B200GEO_ERR_INVALID, "bad argument");
    B200GEO_CUDA(cudaSetDevice(g->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    void *p = 0;
    B200GEO_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    g->peer_buf[side][which] = (char *)p;
    return B200GEO_OK;
}

int b200geo_halo_push(b200geo_grid *g, int side, int width, void *stream)
{
    if (!g || (side != 0 && side != 1)) return fail(B200GEO_ERR_INVALID, "bad argument");
    const int a = g->slab_axis, gz = g->g[a], nz = g->d[a];
    if (width < 1 || width > gz || width > nz) return fail(B200GEO_ERR_INVALID, "bad halo width");
    // both ranks step in lock step, so the neighbour's current buffer has the same index as ours
    char *peer = g->peer_buf[side][g->cur];
    if (!peer) return fail(B200GEO_ERR_LOGIC, "no peer buffer opened on this side");
    B200GEO_CUDA(cudaSetDevice(g->device));
    for (int m = 0; m < g->n; ++m) {
        const MemberLayout& L = g->m[m];
        int64_t src_plane = side == 0 ? gz : gz + nz - width;
        // our low-side boundary lands in the neighbour's high-side ghost planes and vice versa
        // (equal slab thickness on both sides is assumed for the peer's layout)
        int64_t dst_plane = side == 0 ? gz + nz : gz - width;
        int64_t slice = g->slice_elems(m);
        size_t bytes = (size_t)width * slice * L.elem;
        B200GEO_CUDA(cudaMemcpyAsync(peer + L.offset + dst_plane * slice * L.elem,
                                     g->buf[g->cur] + L.offset + src_plane * slice * L.elem,
                                     bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    }
    return B200GEO_OK;
}

int b200geo_stats_enable(b200geo_grid *g, int on)
{
    if (!g) return fail(B200GEO_ERR_INVALID, "null grid");
    g->stats_on = on != 0;
    return B200GEO_OK;
}

int b200geo_stats(b200geo_grid *g, double out[3])
{
    if (!g || !out) return fail(B200GEO_ERR_INVALID, "null argument");
    out[0] = g->t_update;
    out[1] = g->t_ghost;
    out[2] = (double)g->sweeps;
    return B200GEO_OK;
}

}
