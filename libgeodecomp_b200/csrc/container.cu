// ContainerCell grids: ID-keyed cargo (meshfree / unstructured models on the regular container grid).
//
// Reference path replaced (paths relative to /root/reference/src/libgeodecomp/):
//   ContainerCell::update = copyOver + updateCargo                 storage/containercell.h:170-200
//   NeighborhoodAdapter::operator[](id)                            storage/neighborhoodadapter.h:45-65
//   ContainerCell::operator[](id) (upper_bound over ascending ids) storage/containercell.h:107-121
//   the cargo's update (mesh element of the Voronoi example)       src/examples/voronoi/main.cpp:41-54
//
// B200 design. On the CPU every cargo update searches up to 3^DIM containers per neighbour ID, every step. The
// reference never adds or removes cargo while a simulation runs (containercell.h:44-46), so here the search runs ONCE
// after a load (resolve_kernel: the reference's search order, first hit wins) and leaves a link table; a sweep is then
// a pure gather, bound by the HBM traffic of the link table:
//   slot store    the interchange format, container by container (what load / save / the resolver see)
//   compact store one entry per LIVE cargo (exclusive scan of the counts): value[2] (double buffered), influx,
//                 neighbour count, and link[j][i] = compact index of cargo i's j-th neighbour, j-major, so the 32
//                 lanes of a warp read 32 consecutive links per j and the gathered values sit close together in L2.
// The cargo of the edge container (found by lookups beyond a Cube boundary) sits behind the interior cargo in the
// compact store and is never updated. The sum runs in list order from 0.0 with plain adds and one IEEE division, the
// expression tree of the model, hence bit-identical results (the TU is built -fmad=false like the other parity TUs).
#include <cub/device/device_scan.cuh>

#include <cstring>
#include <new>
#include <string>

#include "grid.h"

using namespace b200geo;

struct b200geo_containergrid {
    b200geo_containergrid_desc desc;
    int device;
    int d[3];
    int ndims;                // the Moore box has 3^ndims containers
    int cap, maxnb;
    int64_t ncont;            // interior containers; container `ncont` is the edge container
    // slot store (ncont + 1 containers)
    int32_t *counts;          // [ncont + 2]: one trailing zero so that the exclusive scan ends with the total
    int32_t *ids;
    double *values;
    double *influx;
    int32_t *nbc;
    int32_t *nbids;
    // compact store
    int32_t *offsets;         // [ncont + 2]
    double *cval[2];
    double *cinflux;
    int32_t *cnbc;
    int32_t *link;            // [maxnb][stride]
    int64_t n_interior, n_total, stride, n_links;
    int cur;
    bool dirty;               // slot store changed since the links were resolved
    bool slot_values_stale;   // sweeps ran since the slot values were written
    int32_t *err;             // device: [0] id not found, [1] that id, [2] count out of range, [3] ids do not ascend
    void *scan_tmp;
    size_t scan_tmp_bytes;
    char *staging;
    size_t staging_bytes;
    uint64_t rebuilds, sweeps;
};

namespace {

struct Dims {
    int d[3];
    int wrap[3];
    int ndims;
};

// one array of the interchange format <-> the slot store, for a box of containers; `per` elements per container,
// `per_slot` elements per cargo slot. Saving writes zeros for slots >= count.
template<typename T>
__global__ void box_copy_kernel(T *store, T *buf, const int32_t *counts, int per, int per_slot, int ox, int oy, int oz,
                                int bx, int by, int64_t total, int nx, int ny, int save)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    int64_t b = t / per;
    int e = (int)(t - b * per);
    int x = (int)(b % bx);
    int y = (int)((b / bx) % by);
    int z = (int)(b / ((int64_t)bx * by));
    int64_t c = ((int64_t)(z + oz) * ny + (y + oy)) * nx + (x + ox);
    if (save) {
        T v = store[c * per + e];
        if (per_slot > 0 && e / per_slot >= counts[c]) v = T(0);
        buf[t] = v;
    } else {
        store[c * per + e] = buf[t];
    }
}

// slot store -> compact store (all containers, the edge container included); validates counts and id order
__global__ void compact_kernel(const int32_t *counts, const int32_t *offsets, const int32_t *ids, const double *values,
                               const double *influx, const int32_t *nbc, double *cval0, double *cval1, double *cinflux,
                               int32_t *cnbc, int cap, int maxnb, int64_t slots, int32_t *err)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= slots) return;
    int64_t c = t / cap;
    int s = (int)(t - c * cap);
    int n = counts[c];
    if (s == 0 && (n < 0 || n > cap)) err[2] = 1;
    if (s >= n || n > cap) return;
    if (s > 0 && ids[t - 1] >= ids[t]) err[3] = 1;
    int k = nbc[t];
    if (k < 0 || k > maxnb) {
        err[2] = 1;
        k = 0;
    }
    int64_t i = (int64_t)offsets[c] + s;
    double v = values[t];
    cval0[i] = v;
    cval1[i] = v;
    cinflux[i] = influx[t];
    cnbc[i] = k;
}

// ContainerCell::operator[](id): upper_bound, then look at the element before it (containercell.h:107-121)
__device__ __forceinline__ int find_id(const int32_t *ids, int n, int32_t id)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (id < ids[mid]) hi = mid;
        else lo = mid + 1;
    }
    return (lo > 0 && ids[lo - 1] == id) ? lo - 1 : -1;
}

// NeighborhoodAdapter::operator[] for every neighbour ID of every interior cargo, once
__global__ void resolve_kernel(const int32_t *counts, const int32_t *offsets, const int32_t *ids, const int32_t *nbc,
                               const int32_t *nbids, int32_t *link, int64_t stride, int cap, int maxnb, int64_t ncont,
                               Dims dims, int32_t *err)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ncont * cap) return;
    int64_t c = t / cap;
    int s = (int)(t - c * cap);
    if (s >= counts[c]) return;
    int k = nbc[t];
    if (k < 0 || k > maxnb) return;
    int cx = (int)(c % dims.d[0]);
    int cy = (int)((c / dims.d[0]) % dims.d[1]);
    int cz = (int)(c / ((int64_t)dims.d[0] * dims.d[1]));
    int64_t i = (int64_t)offsets[c] + s;
    int zlo = dims.ndims == 3 ? -1 : 0, zhi = dims.ndims == 3 ? 1 : 0;
    for (int j = 0; j < k; ++j) {
        int32_t id = nbids[t * maxnb + j];
        int64_t found = -1;
        int pos = find_id(ids + c * cap, min(counts[c], cap), id);
        if (pos >= 0) found = (int64_t)offsets[c] + pos;
        for (int dz = zlo; dz <= zhi && found < 0; ++dz) {
            for (int dy = -1; dy <= 1 && found < 0; ++dy) {
                for (int dx = -1; dx <= 1 && found < 0; ++dx) {
                    if (dx == 0 && dy == 0 && dz == 0) continue;
                    int p[3] = {cx + dx, cy + dy, cz + dz};
                    bool outside = false;
                    for (int a = 0; a < 3; ++a) {
                        if (p[a] < 0 || p[a] >= dims.d[a]) {
                            if (dims.wrap[a]) p[a] = (p[a] + dims.d[a]) % dims.d[a];
                            else outside = true;
                        }
                    }
                    int64_t o = outside ? ncont : ((int64_t)p[2] * dims.d[1] + p[1]) * dims.d[0] + p[0];
                    pos = find_id(ids + o * cap, min(counts[o], cap), id);
                    if (pos >= 0) found = (int64_t)offsets[o] + pos;
                }
            }
        }
        if (found < 0) {
            if (atomicExch(&err[0], 1) == 0) err[1] = id;
            found = i;
        }
        link[(int64_t)j * stride + i] = (int32_t)found;
    }
}

// one sweep: the cargo's update() against the old values
__global__ void __launch_bounds__(256) sweep_kernel(const double *__restrict__ old_val, double *__restrict__ new_val,
                                                     const double *__restrict__ influx, const int32_t *__restrict__ nbc,
                                                     const int32_t *__restrict__ link, int64_t stride, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int k = __ldg(nbc + i);
    const int32_t *l = link + i;
    double t = 0.0;
    int j = 0;
    for (; j + 4 <= k; j += 4) {
        int32_t a = __ldg(l + (int64_t)j * stride);
        int32_t b = __ldg(l + (int64_t)(j + 1) * stride);
        int32_t c = __ldg(l + (int64_t)(j + 2) * stride);
        int32_t d = __ldg(l + (int64_t)(j + 3) * stride);
        double va = old_val[a], vb = old_val[b], vc = old_val[c], vd = old_val[d];
        t += va;
        t += vb;
        t += vc;
        t += vd;
    }
    for (; j < k; ++j) t += old_val[__ldg(l + (int64_t)j * stride)];
    // temperature / neighborIDs.size(): the size_t converts to double (voronoi/main.cpp:53)
    new_val[i] = __ldg(influx + i) + t / (double)(unsigned long long)k;
}

// compact values -> slot store (interior containers)
__global__ void scatter_kernel(const int32_t *counts, const int32_t *offsets, const double *cval, double *values, int cap,
                               int64_t slots)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= slots) return;
    int64_t c = t / cap;
    int s = (int)(t - c * cap);
    if (s < counts[c]) values[t] = cval[(int64_t)offsets[c] + s];
}

inline unsigned blocks_for(int64_t n, int threads = 256)
{
    return (unsigned)((n + threads - 1) / threads);
}

int ensure_staging(b200geo_containergrid *g, size_t bytes)
{
    if (bytes <= g->staging_bytes) return B200GEO_OK;
    if (g->staging) cudaFree(g->staging);
    g->staging = 0;
    g->staging_bytes = 0;
    cudaError_t e = cudaMalloc((void **)&g->staging, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(B200GEO_ERR_NOMEM, std::string("cudaMalloc failed: ") + cudaGetErrorString(e));
    }
    g->staging_bytes = bytes;
    return B200GEO_OK;
}

bool valid_box(const b200geo_containergrid *g, const int32_t origin[3], const int32_t dim[3])
{
    for (int i = 0; i < 3; ++i)
        if (origin[i] < 0 || dim[i] < 1 || origin[i] + dim[i] > g->d[i]) return false;
    return true;
}

int scatter_values(b200geo_containergrid *g, cudaStream_t s)
{
    if (!g->slot_values_stale) return B200GEO_OK;
    int64_t slots = g->ncont * g->cap;
    scatter_kernel<<<blocks_for(slots), 256, 0, s>>>(g->counts, g->offsets, g->cval[g->cur], g->values, g->cap, slots);
    count_launch();
    B200GEO_CUDA(cudaGetLastError());
    g->slot_values_stale = false;
    return B200GEO_OK;
}

// one array of a box, either direction; `first` = container the box starts at when it is the edge container (-1: a box
// of interior containers)
template<typename T>
int box_array(b200geo_containergrid *g, T *store, T *user, int per, int per_slot, const int32_t origin[3], const int32_t dim[3],
              int64_t edge_container, int location, bool save, cudaStream_t s)
{
    if (!user) return B200GEO_OK;
    int64_t cells = (int64_t)dim[0] * dim[1] * dim[2];
    int64_t total = cells * per;
    size_t bytes = (size_t)total * sizeof(T);
    T *dev = user;
    if (location == B200GEO_HOST) {
        int rc = ensure_staging(g, bytes);
        if (rc) return rc;
        dev = (T *)g->staging;
        if (!save) B200GEO_CUDA(cudaMemcpyAsync(dev, user, bytes, cudaMemcpyHostToDevice, s));
    }
    if (edge_container >= 0) {
        // the edge container is one contiguous record of the slot store
        box_copy_kernel<T><<<blocks_for(total), 256, 0, s>>>(store + edge_container * per, dev, g->counts + edge_container, per,
                                                             per_slot, 0, 0, 0, 1, 1, total, 1, 1, save ? 1 : 0);
    } else {
        box_copy_kernel<T><<<blocks_for(total), 256, 0, s>>>(store, dev, g->counts, per, per_slot, origin[0], origin[1], origin[2],
                                                             dim[0], dim[1], total, g->d[0], g->d[1], save ? 1 : 0);
    }
    count_launch();
    B200GEO_CUDA(cudaGetLastError());
    if (location == B200GEO_HOST) {
        if (save) B200GEO_CUDA(cudaMemcpyAsync(user, dev, bytes, cudaMemcpyDeviceToHost, s));
        // the staging buffer is reused by the next array
        B200GEO_CUDA(cudaStreamSynchronize(s));
    }
    return B200GEO_OK;
}

int box_io(b200geo_containergrid *g, const int32_t origin[3], const int32_t dim[3], const b200geo_container_box *box,
           int64_t edge_container, int location, bool save, cudaStream_t s)
{
    int rc;
    // counts first: the other arrays are masked with them when saving
    if ((rc = box_array<int32_t>(g, g->counts, box->counts, 1, 0, origin, dim, edge_container, location, save, s))) return rc;
    if ((rc = box_array<int32_t>(g, g->ids, box->ids, g->cap, 1, origin, dim, edge_container, location, save, s))) return rc;
    if ((rc = box_array<double>(g, g->values, box->values, g->cap, 1, origin, dim, edge_container, location, save, s))) return rc;
    if ((rc = box_array<double>(g, g->influx, box->influx, g->cap, 1, origin, dim, edge_container, location, save, s))) return rc;
    if ((rc = box_array<int32_t>(g, g->nbc, box->nb_counts, g->cap, 1, origin, dim, edge_container, location, save, s))) return rc;
    if ((rc = box_array<int32_t>(g, g->nbids, box->nb_ids, g->cap * g->maxnb, g->maxnb, origin, dim, edge_container, location, save, s)))
        return rc;
    if (save) B200GEO_CUDA(cudaStreamSynchronize(s));
    return B200GEO_OK;
}

void free_compact(b200geo_containergrid *g)
{
    for (int b = 0; b < 2; ++b) {
        if (g->cval[b]) cudaFree(g->cval[b]);
        g->cval[b] = 0;
    }
    if (g->cinflux) cudaFree(g->cinflux);
    if (g->cnbc) cudaFree(g->cnbc);
    if (g->link) cudaFree(g->link);
    g->cinflux = 0;
    g->cnbc = 0;
    g->link = 0;
}

// scan the counts, rebuild the compact store, resolve the links
int rebuild(b200geo_containergrid *g, cudaStream_t s)
{
    int64_t n_scan = g->ncont + 2;
    B200GEO_CUDA(cudaMemsetAsync(g->err, 0, 4 * sizeof(int32_t), s));
    size_t need = 0;
    B200GEO_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, need, g->counts, g->offsets, (int)n_scan, s));
    if (need > g->scan_tmp_bytes) {
        if (g->scan_tmp) cudaFree(g->scan_tmp);
        g->scan_tmp = 0;
        g->scan_tmp_bytes = 0;
        cudaError_t e = cudaMalloc(&g->scan_tmp, need);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(B200GEO_ERR_NOMEM, std::string("cudaMalloc failed: ") + cudaGetErrorString(e));
        }
        g->scan_tmp_bytes = need;
    }
    B200GEO_CUDA(cub::DeviceScan::ExclusiveSum(g->scan_tmp, need, g->counts, g->offsets, (int)n_scan, s));
    count_launch();
    int32_t tail[2];
    B200GEO_CUDA(cudaMemcpyAsync(tail, g->offsets + g->ncont, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    B200GEO_CUDA(cudaStreamSynchronize(s));
    int64_t slots_all = (g->ncont + 1) * g->cap;
    if (tail[0] < 0 || tail[1] < tail[0] || tail[1] > slots_all)
        return fail(B200GEO_ERR_OUT_OF_RANGE, "ContainerCell capacity exeeded");
    free_compact(g);
    g->n_interior = tail[0];
    g->n_total = tail[1];
    g->stride = (g->n_interior + 31) / 32 * 32;
    size_t nt = (size_t)(g->n_total > 0 ? g->n_total : 1);
    size_t nl = (size_t)(g->stride > 0 ? g->stride : 32) * g->maxnb;
    cudaError_t e = cudaSuccess;
    for (int b = 0; b < 2 && e == cudaSuccess; ++b) e = cudaMalloc((void **)&g->cval[b], nt * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc((void **)&g->cinflux, nt * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc((void **)&g->cnbc, nt * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMalloc((void **)&g->link, nl * sizeof(int32_t));
    if (e != cudaSuccess) {
        free_compact(g);
        cudaGetLastError();
        return fail(B200GEO_ERR_NOMEM, std::string("cudaMalloc failed: ") + cudaGetErrorString(e));
    }
    compact_kernel<<<blocks_for(slots_all), 256, 0, s>>>(g->counts, g->offsets, g->ids, g->values, g->influx, g->nbc, g->cval[0],
                                                         g->cval[1], g->cinflux, g->cnbc, g->cap, g->maxnb, slots_all, g->err);
    count_launch();
    B200GEO_CUDA(cudaGetLastError());
    Dims dims;
    for (int i = 0; i < 3; ++i) {
        dims.d[i] = g->d[i];
        dims.wrap[i] = g->desc.ghost_mode[i][0] == B200GEO_GHOST_WRAP;
    }
    dims.ndims = g->ndims;
    int64_t slots = g->ncont * g->cap;
    resolve_kernel<<<blocks_for(slots, 128), 128, 0, s>>>(g->counts, g->offsets, g->ids, g->nbc, g->nbids, g->link, g->stride, g->cap,
                                                          g->maxnb, g->ncont, dims, g->err);
    count_launch();
    B200GEO_CUDA(cudaGetLastError());
    int32_t err[4];
    B200GEO_CUDA(cudaMemcpyAsync(err, g->err, sizeof(err), cudaMemcpyDeviceToHost, s));
    B200GEO_CUDA(cudaStreamSynchronize(s));
    if (err[2]) return fail(B200GEO_ERR_OUT_OF_RANGE, "ContainerCell capacity exeeded");
    if (err[3]) return fail(B200GEO_ERR_INVALID, "the ids of a container must ascend (ContainerCell::insert keeps them sorted)");
    if (err[0]) return fail(B200GEO_ERR_LOGIC, "id not found: could not find id " + std::to_string(err[1]) + " in neighborhood");
    // links of the interior cargo
    int32_t *knb = new (std::nothrow) int32_t[(size_t)(g->n_interior > 0 ? g->n_interior : 1)];
    g->n_links = 0;
    if (knb) {
        if (g->n_interior > 0 &&
            cudaMemcpy(knb, g->cnbc, (size_t)g->n_interior * sizeof(int32_t), cudaMemcpyDeviceToHost) == cudaSuccess) {
            for (int64_t i = 0; i < g->n_interior; ++i) g->n_links += knb[i];
        }
        delete[] knb;
    }
    g->cur = 0;
    g->dirty = false;
    g->slot_values_stale = false;
    ++g->rebuilds;
    return B200GEO_OK;
}

bool box_complete(const b200geo_container_box *b)
{
    return b && b->counts && b->ids && b->values && b->influx && b->nb_counts && b->nb_ids;
}

}

extern "C" {

int b200geo_containergrid_create(const b200geo_containergrid_desc *desc, int device, b200geo_containergrid **out)
{
    if (!desc || !out) return fail(B200GEO_ERR_INVALID, "null argument");
    for (int i = 0; i < 3; ++i) {
        if (desc->dim[i] < 1) return fail(B200GEO_ERR_INVALID, "grid dimension must be >= 1");
        int lo = desc->ghost_mode[i][0], hi = desc->ghost_mode[i][1];
        if (lo == B200GEO_GHOST_PEER || hi == B200GEO_GHOST_PEER)
            return fail(B200GEO_ERR_LOGIC, "ContainerCell grids live on one device: no PEER ghost layers");
        if ((lo != B200GEO_GHOST_EDGE && lo != B200GEO_GHOST_WRAP) || lo != hi)
            return fail(B200GEO_ERR_INVALID, "ghost mode must be EDGE or WRAP, alike on both sides of an axis");
    }
    if (desc->n_dims != 2 && desc->n_dims != 3) return fail(B200GEO_ERR_INVALID, "n_dims must be 2 or 3");
    if (desc->n_dims == 2 && desc->dim[2] != 1) return fail(B200GEO_ERR_INVALID, "a 2-D grid has dim[2] = 1");
    if (desc->capacity < 1 || desc->capacity > 4096) return fail(B200GEO_ERR_INVALID, "capacity must be 1..4096");
    if (desc->max_neighbors < 1 || desc->max_neighbors > 64) return fail(B200GEO_ERR_INVALID, "max_neighbors must be 1..64");
    int64_t ncont = (int64_t)desc->dim[0] * desc->dim[1] * desc->dim[2];
    if ((ncont + 1) * desc->capacity >= ((int64_t)1 << 31))
        return fail(B200GEO_ERR_OUT_OF_RANGE, "more than 2^31 cargo slots");
    B200GEO_CUDA(cudaSetDevice(device));
    b200geo_containergrid *g = new (std::nothrow) b200geo_containergrid();
    if (!g) return fail(B200GEO_ERR_NOMEM, "out of host memory");
    memset(g, 0, sizeof(*g));
    g->desc = *desc;
    g->device = device;
    for (int i = 0; i < 3; ++i) g->d[i] = desc->dim[i];
    g->ndims = desc->n_dims;
    g->cap = desc->capacity;
    g->maxnb = desc->max_neighbors;
    g->ncont = ncont;
    size_t slots = (size_t)(ncont + 1) * g->cap;
    cudaError_t e = cudaMalloc((void **)&g->counts, (size_t)(ncont + 2) * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMalloc((void **)&g->offsets, (size_t)(ncont + 2) * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMalloc((void **)&g->ids, slots * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMalloc((void **)&g->values, slots * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc((void **)&g->influx, slots * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc((void **)&g->nbc, slots * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMalloc((void **)&g->nbids, slots * g->maxnb * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMalloc((void **)&g->err, 4 * sizeof(int32_t));
    if (e != cudaSuccess) {
        b200geo_containergrid_destroy(g);
        cudaGetLastError();
        return fail(B200GEO_ERR_NOMEM, std::string("cudaMalloc failed: ") + cudaGetErrorString(e));
    }
    // every container starts empty, like ContainerCell() (containercell.h:58-60)
    cudaMemset(g->counts, 0, (size_t)(ncont + 2) * sizeof(int32_t));
    cudaMemset(g->ids, 0, slots * sizeof(int32_t));
    cudaMemset(g->values, 0, slots * sizeof(double));
    cudaMemset(g->influx, 0, slots * sizeof(double));
    cudaMemset(g->nbc, 0, slots * sizeof(int32_t));
    cudaMemset(g->nbids, 0, slots * g->maxnb * sizeof(int32_t));
    g->dirty = true;
    *out = g;
    return B200GEO_OK;
}

int b200geo_containergrid_destroy(b200geo_containergrid *g)
{
    if (!g) return B200GEO_OK;
    cudaSetDevice(g->device);
    free_compact(g);
    if (g->counts) cudaFree(g->counts);
    if (g->offsets) cudaFree(g->offsets);
    if (g->ids) cudaFree(g->ids);
    if (g->values) cudaFree(g->values);
    if (g->influx) cudaFree(g->influx);
    if (g->nbc) cudaFree(g->nbc);
    if (g->nbids) cudaFree(g->nbids);
    if (g->err) cudaFree(g->err);
    if (g->scan_tmp) cudaFree(g->scan_tmp);
    if (g->staging) cudaFree(g->staging);
    delete g;
    return B200GEO_OK;
}

int b200geo_containergrid_load(b200geo_containergrid *g, const int32_t origin[3], const int32_t dim[3],
                               const b200geo_container_box *box, int location, void *stream)
{
    if (!g || !origin || !dim) return fail(B200GEO_ERR_INVALID, "null argument");
    if (!box_complete(box)) return fail(B200GEO_ERR_INVALID, "a load needs every array of the box");
    if (!valid_box(g, origin, dim)) return fail(B200GEO_ERR_INVALID, "box outside the grid");
    if (location != B200GEO_HOST && location != B200GEO_CUDA_DEVICE) return fail(B200GEO_ERR_INVALID, "bad memory location");
    B200GEO_CUDA(cudaSetDevice(g->device));
    cudaStream_t s = (cudaStream_t)stream;
    int rc = scatter_values(g, s);
    if (rc) return rc;
    g->dirty = true;
    return box_io(g, origin, dim, box, -1, location, false, s);
}

int b200geo_containergrid_save(const b200geo_containergrid *cg, const int32_t origin[3], const int32_t dim[3],
                               const b200geo_container_box *box, int location, void *stream)
{
    if (!cg || !origin || !dim || !box) return fail(B200GEO_ERR_INVALID, "null argument");
    b200geo_containergrid *g = const_cast<b200geo_containergrid *>(cg);
    if (!valid_box(g, origin, dim)) return fail(B200GEO_ERR_INVALID, "box outside the grid");
    if (location != B200GEO_HOST && location != B200GEO_CUDA_DEVICE) return fail(B200GEO_ERR_INVALID, "bad memory location");
    B200GEO_CUDA(cudaSetDevice(g->device));
    cudaStream_t s = (cudaStream_t)stream;
    int rc = scatter_values(g, s);
    if (rc) return rc;
    return box_io(g, origin, dim, box, -1, location, true, s);
}

int b200geo_containergrid_set_edge(b200geo_containergrid *g, const b200geo_container_box *cell)
{
    if (!g) return fail(B200GEO_ERR_INVALID, "null argument");
    if (!box_complete(cell)) return fail(B200GEO_ERR_INVALID, "the edge container needs every array");
    if (cell->counts[0] < 0 || cell->counts[0] > g->cap) return fail(B200GEO_ERR_OUT_OF_RANGE, "ContainerCell capacity exeeded");
    B200GEO_CUDA(cudaSetDevice(g->device));
    int rc = scatter_values(g, 0);
    if (rc) return rc;
    g->dirty = true;
    const int32_t o[3] = {0, 0, 0}, d[3] = {1, 1, 1};
    return box_io(g, o, d, cell, g->ncont, B200GEO_HOST, false, 0);
}

int b200geo_containergrid_get_edge(const b200geo_containergrid *cg, const b200geo_container_box *cell)
{
    if (!cg || !cell) return fail(B200GEO_ERR_INVALID, "null argument");
    b200geo_containergrid *g = const_cast<b200geo_containergrid *>(cg);
    B200GEO_CUDA(cudaSetDevice(g->device));
    const int32_t o[3] = {0, 0, 0}, d[3] = {1, 1, 1};
    return box_io(g, o, d, cell, g->ncont, B200GEO_HOST, true, 0);
}

int b200geo_containergrid_step(b200geo_containergrid *g, uint32_t first_nano_step, uint32_t n_steps, void *stream)
{
    (void)first_nano_step;   // the bound cargo's update() does not look at the nano step
    if (!g) return fail(B200GEO_ERR_INVALID, "null argument");
    if (n_steps == 0) return B200GEO_OK;
    B200GEO_CUDA(cudaSetDevice(g->device));
    cudaStream_t s = (cudaStream_t)stream;
    if (g->dirty) {
        int rc = rebuild(g, s);
        if (rc) return rc;
    }
    if (g->n_interior > 0) {
        for (uint32_t t = 0; t < n_steps; ++t) {
            sweep_kernel<<<blocks_for(g->n_interior), 256, 0, s>>>(g->cval[g->cur], g->cval[g->cur ^ 1], g->cinflux, g->cnbc, g->link,
                                                                   g->stride, g->n_interior);
            g->cur ^= 1;
        }
        count_launch(n_steps);
        B200GEO_CUDA(cudaGetLastError());
        g->slot_values_stale = true;
    }
    g->sweeps += n_steps;
    return B200GEO_OK;
}

int b200geo_containergrid_stats(const b200geo_containergrid *g, uint64_t out[4])
{
    if (!g || !out) return fail(B200GEO_ERR_INVALID, "null argument");
    out[0] = (uint64_t)g->n_interior;
    out[1] = (uint64_t)g->n_links;
    out[2] = g->rebuilds;
    out[3] = g->sweeps;
    return B200GEO_OK;
}

}
