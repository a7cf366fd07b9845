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
//   compact store one entry per LIVE cargo (exclusive scan of the counts), in the sliced-ELLPACK order sparse
//                 matrix-vector kernels use (SELL-32-1024): inside every window of 1024 consecutive cargo items the
//                 items are sorted by falling neighbour count, and every chunk of 32 items (one warp) stores its
//                 links j-major with the chunk's own width: link[chunk_off + j * 32 + lane]. The 32 lanes of a warp
//                 read 32 consecutive links per j, all lanes of a warp loop (almost) equally long, and the table
//                 holds hardly more than the links there are; the first run of this path kept maxnb x stride links
//                 and fetched 1.74 x the link bytes it used (profiles/r6a). Windows keep the gathered values close
//                 together in L2. Per entry: value[2] (double buffered), influx, neighbour count.
// The cargo of the edge container (found by lookups beyond a Cube boundary) sits behind the interior windows in the
// compact store and is never updated. The sum runs in list order from 0.0 with plain adds and one IEEE division, the
// expression tree of the model, hence bit-identical results (the TU is built -fmad=false like the other parity TUs).
#include <cub/block/block_radix_sort.cuh>
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
    int32_t *offsets;         // [ncont + 2]: first compact index of a container (container order)
    int32_t *pos;             // [n_interior]: compact index -> position in the sorted windows
    int32_t *deg;             // [n_interior]: neighbour count by compact index (sort key)
    double *cval[2];          // by position; the edge container's cargo at n_padded + slot
    double *cinflux;
    int32_t *cnbc;            // by position: neighbour count (0 for the padding of the last window)
    int64_t *chunk_off;       // [n_padded / 32 + 1]: first link of a chunk of 32 positions
    int32_t *link;            // chunk by chunk: [width of the chunk][32]
    unsigned long long *link_count;   // device: sum of the neighbour counts
    int64_t n_interior, n_total, n_padded, n_links, n_link_slots;
    // tile layout ("container.kernel" = 1): a tile = tile_g containers along x; its cargo sorted by neighbour count
    // into tile_w positions; values stay in container order (cval by compact index), links are 16-bit indices into
    // the tile's staged neighbourhood ((rows x (tile_g + 2) containers) x capacity slots in shared memory)
    int kernel;               // layout the compact store was built for
    int tile_g, tile_w;
    int64_t n_tiles, tiles_per_row;
    int32_t *ptarget;         // [n_padded]: compact index of the cargo at a position, -1 for padding
    uint16_t *link16;
    size_t tile_smem;
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

const int WINDOW = 1024;   // sigma of SELL-C-sigma: cargo items sorted by neighbour count inside windows of this many
const int CHUNK = 32;      // C: one warp
const int MAX_STAGED = 9 * 34;   // containers a tile stages at most: 3 x 3 rows of (32 + 2)

// every count within 0..capacity? (before anything is indexed with their prefix sums)
__global__ void validate_counts_kernel(const int32_t *counts, int64_t n, int cap, int32_t *err)
{
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n && (counts[c] < 0 || counts[c] > cap)) err[2] = 1;
}

// neighbour count of every interior cargo item by compact index; validates counts, id order and neighbour counts
__global__ void degree_kernel(const int32_t *counts, const int32_t *offsets, const int32_t *ids, const int32_t *nbc, int32_t *deg,
                              int cap, int maxnb, int64_t ncont, unsigned long long *link_count, int32_t *err)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (ncont + 1) * cap) return;
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
    if (c < ncont) {
        deg[(int64_t)offsets[c] + s] = k;
        if (k) atomicAdd(link_count, (unsigned long long)k);
    }
}

// one CTA per window: sort the window's cargo by falling neighbour count (the padding of the last window goes last)
__global__ void __launch_bounds__(256) window_sort_kernel(const int32_t *deg, int32_t *pos, int32_t *cnbc, int maxnb, int64_t n)
{
    typedef cub::BlockRadixSort<int, 256, WINDOW / 256, int> Sort;
    __shared__ typename Sort::TempStorage tmp;
    int keys[WINDOW / 256], vals[WINDOW / 256];
    int64_t base = (int64_t)blockIdx.x * WINDOW;
#pragma unroll
    for (int e = 0; e < WINDOW / 256; ++e) {
        int64_t i = base + threadIdx.x * (WINDOW / 256) + e;
        keys[e] = i < n ? maxnb - deg[i] : maxnb + 1;
        vals[e] = threadIdx.x * (WINDOW / 256) + e;
    }
    Sort(tmp).Sort(keys, vals, 0, 8);
#pragma unroll
    for (int e = 0; e < WINDOW / 256; ++e) {
        int64_t p = base + threadIdx.x * (WINDOW / 256) + e;
        int64_t i = base + vals[e];
        cnbc[p] = i < n ? maxnb - keys[e] : 0;
        if (i < n) pos[i] = (int32_t)p;
    }
}

// links a chunk of 32 positions stores: 32 x the neighbour count of its first (= longest) item
__global__ void chunk_width_kernel(const int32_t *cnbc, int64_t *chunk_links, int64_t chunks)
{
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c <= chunks) chunk_links[c] = c < chunks ? CHUNK * cnbc[c * CHUNK] : 0;
}

// slot store -> compact store (all containers, the edge container included)
__global__ void compact_kernel(const int32_t *counts, const int32_t *offsets, const int32_t *pos, const double *values,
                               const double *influx, double *cval0, double *cval1, double *cinflux, int cap, int64_t ncont,
                               int64_t n_interior, int64_t n_padded)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (ncont + 1) * cap) return;
    int64_t c = t / cap;
    int s = (int)(t - c * cap);
    if (s >= counts[c]) return;
    int64_t i = (int64_t)offsets[c] + s;
    int64_t p = c < ncont ? pos[i] : n_padded + (i - n_interior);
    double v = values[t];
    cval0[p] = v;
    cval1[p] = v;
    cinflux[p] = influx[t];
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
__global__ void resolve_kernel(const int32_t *counts, const int32_t *offsets, const int32_t *pos, const int32_t *ids,
                               const int32_t *nbc, const int32_t *nbids, const int64_t *chunk_off, int32_t *link, int cap, int maxnb,
                               int64_t ncont, int64_t n_interior, int64_t n_padded, Dims dims, int32_t *err)
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
    int64_t p = pos[(int64_t)offsets[c] + s];
    int32_t *out = link + chunk_off[p / CHUNK] + (p % CHUNK);
    int zlo = dims.ndims == 3 ? -1 : 0, zhi = dims.ndims == 3 ? 1 : 0;
    for (int j = 0; j < k; ++j) {
        int32_t id = nbids[t * maxnb + j];
        int64_t found = -1;
        int at = find_id(ids + c * cap, min(counts[c], cap), id);
        if (at >= 0) found = pos[(int64_t)offsets[c] + at];
        for (int dz = zlo; dz <= zhi && found < 0; ++dz) {
            for (int dy = -1; dy <= 1 && found < 0; ++dy) {
                for (int dx = -1; dx <= 1 && found < 0; ++dx) {
                    if (dx == 0 && dy == 0 && dz == 0) continue;
                    int q[3] = {cx + dx, cy + dy, cz + dz};
                    bool outside = false;
                    for (int a = 0; a < 3; ++a) {
                        if (q[a] < 0 || q[a] >= dims.d[a]) {
                            if (dims.wrap[a]) q[a] = (q[a] + dims.d[a]) % dims.d[a];
                            else outside = true;
                        }
                    }
                    int64_t o = outside ? ncont : ((int64_t)q[2] * dims.d[1] + q[1]) * dims.d[0] + q[0];
                    at = find_id(ids + o * cap, min(counts[o], cap), id);
                    if (at >= 0) found = outside ? n_padded + ((int64_t)offsets[o] + at - n_interior) : pos[(int64_t)offsets[o] + at];
                }
            }
        }
        if (found < 0) {
            if (atomicExch(&err[0], 1) == 0) err[1] = id;
            found = p;
        }
        out[j * CHUNK] = (int32_t)found;
    }
}

// one sweep: the cargo's update() against the old values. One warp per chunk; the links of the next four neighbours
// are on their way while the values of the current four are gathered.
__global__ void __launch_bounds__(256) sweep_kernel(const double *__restrict__ old_val, double *__restrict__ new_val,
                                                     const double *__restrict__ influx, const int32_t *__restrict__ nbc,
                                                     const int64_t *__restrict__ chunk_off, const int32_t *__restrict__ link,
                                                     int64_t n)
{
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    int k = __ldg(nbc + p);
    int64_t chunk = p / CHUNK;
    int64_t first = __ldg(chunk_off + chunk);
    int width = (int)(__ldg(chunk_off + chunk + 1) - first) / CHUNK;   // the same for the 32 lanes of a warp
    const int32_t *l = link + first + (p % CHUNK);
    double fl = __ldg(influx + p);
    double t = 0.0;
    int32_t a[4], b[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) a[u] = u < width ? __ldg(l + u * CHUNK) : 0;
    for (int j = 0; j < width; j += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) b[u] = j + 4 + u < width ? __ldg(l + (j + 4 + u) * CHUNK) : 0;
        double v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = j + u < k ? old_val[a[u]] : 0.0;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (j + u < k) t += v[u];
            a[u] = b[u];
        }
    }
    // temperature / neighborIDs.size(): the size_t converts to double (voronoi/main.cpp:53)
    new_val[p] = fl + t / (double)(unsigned long long)k;
}

// compact values -> slot store (interior containers)
__global__ void scatter_kernel(const int32_t *counts, const int32_t *offsets, const int32_t *pos, const double *cval, double *values,
                               int cap, int64_t slots, int by_position)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= slots) return;
    int64_t c = t / cap;
    int s = (int)(t - c * cap);
    if (s < counts[c]) values[t] = cval[by_position ? pos[(int64_t)offsets[c] + s] : offsets[c] + s];
}

// ---- tile layout: gathers from shared memory ------------------------------------------------------------------------

struct Tiles {
    int g, w, cap, rows, ndims;   // containers per tile, positions per tile, capacity, staged rows (3 or 9)
    int64_t per_row;              // tiles per row of containers
};

// container a tile stages at (staged row r, local x lx), or `ncont` (the edge container) beyond a Cube boundary
__device__ __forceinline__ int64_t staged_container(const Tiles& T, const Dims& D, int64_t ncont, int x0, int cy, int cz, int r, int lx)
{
    int q[3] = {x0 - 1 + lx, cy + r % 3 - 1, T.ndims == 3 ? cz + r / 3 - 1 : cz};
    for (int a = 0; a < 3; ++a) {
        if (q[a] < 0 || q[a] >= D.d[a]) {
            if (!D.wrap[a]) return ncont;
            q[a] = (q[a] % D.d[a] + D.d[a]) % D.d[a];
        }
    }
    return ((int64_t)q[2] * D.d[1] + q[1]) * D.d[0] + q[0];
}

// one CTA per tile: sort the tile's cargo by falling neighbour count into the tile's positions (padding last)
template<int IPT>
__global__ void __launch_bounds__(256) tile_sort_kernel(const int32_t *offsets, const int32_t *deg, int32_t *pos, int32_t *cnbc,
                                                         int32_t *ptarget, int maxnb, Tiles T, int nx)
{
    typedef cub::BlockRadixSort<int, 256, IPT, int> Sort;
    __shared__ typename Sort::TempStorage tmp;
    int64_t tile = blockIdx.x;
    int x0 = (int)(tile % T.per_row) * T.g;
    int64_t c0 = (tile / T.per_row) * nx + x0;
    int gc = min(T.g, nx - x0);
    int start = offsets[c0];
    int n = offsets[c0 + gc] - start;
    int keys[IPT], vals[IPT];
#pragma unroll
    for (int e = 0; e < IPT; ++e) {
        int li = threadIdx.x * IPT + e;
        keys[e] = li < n ? maxnb - deg[start + li] : maxnb + 1;
        vals[e] = li;
    }
    Sort(tmp).Sort(keys, vals, 0, 8);
#pragma unroll
    for (int e = 0; e < IPT; ++e) {
        int64_t p = tile * T.w + threadIdx.x * IPT + e;
        bool real = vals[e] < n;
        cnbc[p] = real ? maxnb - keys[e] : 0;
        ptarget[p] = real ? start + vals[e] : -1;
        if (real) pos[start + vals[e]] = (int32_t)p;
    }
}

// slot store -> values in container order (every container, the edge container last) and the influx by position
__global__ void tile_fill_kernel(const int32_t *counts, const int32_t *offsets, const int32_t *pos, const double *values,
                                 const double *influx, double *cval0, double *cval1, double *cinflux, int cap, int64_t ncont)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (ncont + 1) * cap) return;
    int64_t c = t / cap;
    int s = (int)(t - c * cap);
    if (s >= counts[c]) return;
    int64_t i = (int64_t)offsets[c] + s;
    double v = values[t];
    cval0[i] = v;
    cval1[i] = v;
    if (c < ncont) cinflux[pos[i]] = influx[t];
}

// NeighborhoodAdapter::operator[] once per neighbour ID, as the index of the hit inside the tile's staged neighbourhood
__global__ void tile_resolve_kernel(const int32_t *counts, const int32_t *offsets, const int32_t *pos, const int32_t *ids,
                                    const int32_t *nbc, const int32_t *nbids, const int64_t *chunk_off, uint16_t *link, int maxnb,
                                    int64_t ncont, Tiles T, Dims dims, int32_t *err)
{
    const int cap = T.cap;
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
    int lx0 = cx % T.g + 1;                       // the own container's place in the staged rows
    int own_row = T.ndims == 3 ? 4 : 1;
    int64_t p = pos[(int64_t)offsets[c] + s];
    uint16_t *out = link + chunk_off[p / CHUNK] + (p % CHUNK);
    int zlo = dims.ndims == 3 ? -1 : 0, zhi = dims.ndims == 3 ? 1 : 0;
    for (int j = 0; j < k; ++j) {
        int32_t id = nbids[t * maxnb + j];
        int found = -1;
        int at = find_id(ids + c * cap, min(counts[c], cap), id);
        if (at >= 0) found = (own_row * (T.g + 2) + lx0) * cap + at;
        for (int dz = zlo; dz <= zhi && found < 0; ++dz) {
            for (int dy = -1; dy <= 1 && found < 0; ++dy) {
                for (int dx = -1; dx <= 1 && found < 0; ++dx) {
                    if (dx == 0 && dy == 0 && dz == 0) continue;
                    int q[3] = {cx + dx, cy + dy, cz + dz};
                    bool outside = false;
                    for (int a = 0; a < 3; ++a) {
                        if (q[a] < 0 || q[a] >= dims.d[a]) {
                            if (dims.wrap[a]) q[a] = (q[a] + dims.d[a]) % dims.d[a];
                            else outside = true;
                        }
                    }
                    int64_t o = outside ? ncont : ((int64_t)q[2] * dims.d[1] + q[1]) * dims.d[0] + q[0];
                    at = find_id(ids + o * cap, min(counts[o], cap), id);
                    if (at >= 0) found = (((dz - zlo) * 3 + dy + 1) * (T.g + 2) + lx0 + dx) * cap + at;
                }
            }
        }
        if (found < 0) {
            if (atomicExch(&err[0], 1) == 0) err[1] = id;
            found = 0;
        }
        out[j * CHUNK] = (uint16_t)found;
    }
}

// one sweep, one CTA per tile. Every warp owns CPW chunks of 32 cargo items (positions sorted by falling neighbour
// count): it first puts the loads of their neighbour counts, targets, influx and first four links in flight, then helps
// to stage the old values of the tile's neighbourhood (rows x (g + 2) containers, slot by slot) in shared memory, and
// after the barrier follows its 16-bit links there; the links of the next four neighbours are always on their way.
template<int CPW>
__global__ void __launch_bounds__(256) tile_sweep_kernel(const double *__restrict__ old_val, double *__restrict__ new_val,
                                                          const double *__restrict__ influx, const int32_t *__restrict__ nbc,
                                                          const int32_t *__restrict__ ptarget, const int64_t *__restrict__ chunk_off,
                                                          const uint16_t *__restrict__ link, const int32_t *__restrict__ offsets,
                                                          int64_t ncont, Tiles T, Dims D)
{
    extern __shared__ double sval[];
    const int64_t tile = blockIdx.x;
    const int x0 = (int)(tile % T.per_row) * T.g;
    const int64_t row = tile / T.per_row;
    const int cy = (int)(row % D.d[1]), cz = (int)(row / D.d[1]);
    const int lane = threadIdx.x % CHUNK, warp = threadIdx.x / CHUNK;
    const int64_t c0 = row * D.d[0] + x0;
    const int n_tile = __ldg(offsets + c0 + min(T.g, D.d[0] - x0)) - __ldg(offsets + c0);   // cargo items of this tile

    int k[CPW], target[CPW], width[CPW];
    double fl[CPW];
    const uint16_t *l[CPW];
    uint16_t a[CPW][4];
#pragma unroll
    for (int u = 0; u < CPW; ++u) {
        const int ch = warp * CPW + u;
        const int r = ch * CHUNK + lane;
        const bool real = r < n_tile;               // positions behind the tile's cargo are padding
        const int64_t p = tile * T.w + r;
        k[u] = real ? __ldg(nbc + p) : 0;
        target[u] = real ? __ldg(ptarget + p) : -1;
        fl[u] = real ? __ldg(influx + p) : 0.0;
        int64_t first = 0;
        width[u] = 0;
        if (ch * CHUNK < n_tile) {
            const int64_t chunk = tile * (T.w / CHUNK) + ch;
            first = __ldg(chunk_off + chunk);
            width[u] = (int)(__ldg(chunk_off + chunk + 1) - first) / CHUNK;   // the same for the 32 lanes of a warp
        }
        l[u] = link + first + lane;
#pragma unroll
        for (int v = 0; v < 4; ++v) a[u][v] = v < width[u] ? __ldg(l[u] + v * CHUNK) : (uint16_t)0;
    }

    // where the staged containers' cargo starts and how many items they hold: one thread per container (the index
    // arithmetic of a wrapped neighbourhood costs ~150 instructions; done warp by warp it made the kernel issue-bound,
    // profiles/r6c), then one warp per container copies its values
    __shared__ int s_first[MAX_STAGED], s_count[MAX_STAGED];
    const int staged = T.rows * (T.g + 2);
    for (int q = threadIdx.x; q < staged; q += 256) {
        const int64_t c = staged_container(T, D, ncont, x0, cy, cz, q / (T.g + 2), q % (T.g + 2));
        const int o = __ldg(offsets + c);
        s_first[q] = o;
        s_count[q] = __ldg(offsets + c + 1) - o;
    }
    __syncthreads();
    for (int q = warp; q < staged; q += 8) {
        const int o = s_first[q], n = s_count[q];
        for (int s = lane; s < n; s += CHUNK) sval[q * T.cap + s] = old_val[o + s];
    }
    __syncthreads();

#pragma unroll
    for (int u = 0; u < CPW; ++u) {
        double t = 0.0;
        for (int j = 0; j < width[u]; j += 4) {
            uint16_t b[4];
#pragma unroll
            for (int v = 0; v < 4; ++v) b[v] = j + 4 + v < width[u] ? __ldg(l[u] + (j + 4 + v) * CHUNK) : (uint16_t)0;
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                if (j + v < k[u]) t += sval[a[u][v]];
                a[u][v] = b[v];
            }
        }
        // temperature / neighborIDs.size(): the size_t converts to double (voronoi/main.cpp:53)
        if (target[u] >= 0) new_val[target[u]] = fl[u] + t / (double)(unsigned long long)k[u];
    }
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
    scatter_kernel<<<blocks_for(slots), 256, 0, s>>>(g->counts, g->offsets, g->pos, g->cval[g->cur], g->values, g->cap, slots,
                                                     g->kernel == 0);
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
    if (g->pos) cudaFree(g->pos);
    if (g->deg) cudaFree(g->deg);
    if (g->chunk_off) cudaFree(g->chunk_off);
    if (g->ptarget) cudaFree(g->ptarget);
    if (g->link16) cudaFree(g->link16);
    g->ptarget = 0;
    g->link16 = 0;
    g->cinflux = 0;
    g->cnbc = 0;
    g->link = 0;
    g->pos = 0;
    g->deg = 0;
    g->chunk_off = 0;
}

int ensure_scan_tmp(b200geo_containergrid *g, size_t need)
{
    if (need <= g->scan_tmp_bytes) return B200GEO_OK;
    if (g->scan_tmp) cudaFree(g->scan_tmp);
    g->scan_tmp = 0;
    g->scan_tmp_bytes = 0;
    cudaError_t e = cudaMalloc(&g->scan_tmp, need);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(B200GEO_ERR_NOMEM, std::string("cudaMalloc failed: ") + cudaGetErrorString(e));
    }
    g->scan_tmp_bytes = need;
    return B200GEO_OK;
}

// scan the counts, order the cargo (windows sorted by neighbour count), rebuild the compact store, resolve the links
int rebuild(b200geo_containergrid *g, cudaStream_t s)
{
    int64_t n_scan = g->ncont + 2;
    int rc;
    B200GEO_CUDA(cudaMemsetAsync(g->err, 0, 4 * sizeof(int32_t), s));
    B200GEO_CUDA(cudaMemsetAsync(g->link_count, 0, sizeof(unsigned long long), s));
    validate_counts_kernel<<<blocks_for(g->ncont + 1), 256, 0, s>>>(g->counts, g->ncont + 1, g->cap, g->err);
    count_launch();
    size_t need = 0;
    B200GEO_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, need, g->counts, g->offsets, (int)n_scan, s));
    if ((rc = ensure_scan_tmp(g, need))) return rc;
    B200GEO_CUDA(cub::DeviceScan::ExclusiveSum(g->scan_tmp, need, g->counts, g->offsets, (int)n_scan, s));
    count_launch();
    int32_t tail[2], err[4];
    B200GEO_CUDA(cudaMemcpyAsync(tail, g->offsets + g->ncont, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    B200GEO_CUDA(cudaMemcpyAsync(err, g->err, sizeof(err), cudaMemcpyDeviceToHost, s));
    B200GEO_CUDA(cudaStreamSynchronize(s));
    int64_t slots_all = (g->ncont + 1) * g->cap;
    if (err[2] || tail[0] < 0 || tail[1] < tail[0] || tail[1] > slots_all)
        return fail(B200GEO_ERR_OUT_OF_RANGE, "ContainerCell capacity exeeded");
    free_compact(g);
    g->n_interior = tail[0];
    g->n_total = tail[1];
    // which layout: tiles with staged values where the capacity allows and the tuning asks for it
    g->kernel = 0;
    if (g_tuning.container_kernel == 1 && g->cap <= WINDOW) {
        const int rows = g->ndims == 3 ? 9 : 3;
        int tg = WINDOW / g->cap < 32 ? WINDOW / g->cap : 32;
        while (tg > 1 && (size_t)rows * (tg + 2) * g->cap * sizeof(double) > 96 * 1024) --tg;
        size_t smem = (size_t)rows * (tg + 2) * g->cap * sizeof(double);
        if (tg >= 1 && smem <= 96 * 1024 && (int64_t)rows * (tg + 2) * g->cap <= 65535) {
            g->kernel = 1;
            g->tile_g = tg;
            g->tile_w = (tg * g->cap + 255) / 256 * 256;
            g->tiles_per_row = (g->d[0] + tg - 1) / tg;
            g->n_tiles = g->tiles_per_row * g->d[1] * g->d[2];
            g->tile_smem = smem;
        }
    }
    g->n_padded = g->kernel == 1 ? g->n_tiles * g->tile_w : (g->n_interior + WINDOW - 1) / WINDOW * WINDOW;
    int64_t chunks = g->n_padded / CHUNK;
    // values: by position (windows) with the edge container's cargo behind them, or by compact index (tiles)
    size_t nv = (size_t)((g->kernel == 1 ? g->n_total : g->n_padded + (g->n_total - g->n_interior)) + 1);
    size_t np = (size_t)(g->n_padded + 1);
    size_t ni = (size_t)(g->n_interior > 0 ? g->n_interior : 1);
    cudaError_t e = cudaSuccess;
    for (int b = 0; b < 2 && e == cudaSuccess; ++b) e = cudaMalloc((void **)&g->cval[b], nv * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc((void **)&g->cinflux, (nv > np ? nv : np) * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc((void **)&g->cnbc, np * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMalloc((void **)&g->pos, ni * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMalloc((void **)&g->deg, ni * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMalloc((void **)&g->chunk_off, (size_t)(chunks + 1) * sizeof(int64_t));
    if (e == cudaSuccess && g->kernel == 1) e = cudaMalloc((void **)&g->ptarget, np * sizeof(int32_t));
    if (e != cudaSuccess) {
        free_compact(g);
        cudaGetLastError();
        return fail(B200GEO_ERR_NOMEM, std::string("cudaMalloc failed: ") + cudaGetErrorString(e));
    }
    degree_kernel<<<blocks_for(slots_all), 256, 0, s>>>(g->counts, g->offsets, g->ids, g->nbc, g->deg, g->cap, g->maxnb, g->ncont,
                                                        g->link_count, g->err);
    count_launch();
    B200GEO_CUDA(cudaGetLastError());
    Tiles tiles;
    tiles.g = g->tile_g;
    tiles.w = g->tile_w;
    tiles.cap = g->cap;
    tiles.rows = g->ndims == 3 ? 9 : 3;
    tiles.ndims = g->ndims;
    tiles.per_row = g->tiles_per_row;
    if (g->kernel == 1) {
        unsigned nt = (unsigned)g->n_tiles;
        switch (g->tile_w / 256) {
        case 1: tile_sort_kernel<1><<<nt, 256, 0, s>>>(g->offsets, g->deg, g->pos, g->cnbc, g->ptarget, g->maxnb, tiles, g->d[0]); break;
        case 2: tile_sort_kernel<2><<<nt, 256, 0, s>>>(g->offsets, g->deg, g->pos, g->cnbc, g->ptarget, g->maxnb, tiles, g->d[0]); break;
        case 3: tile_sort_kernel<3><<<nt, 256, 0, s>>>(g->offsets, g->deg, g->pos, g->cnbc, g->ptarget, g->maxnb, tiles, g->d[0]); break;
        default: tile_sort_kernel<4><<<nt, 256, 0, s>>>(g->offsets, g->deg, g->pos, g->cnbc, g->ptarget, g->maxnb, tiles, g->d[0]); break;
        }
        count_launch();
        B200GEO_CUDA(cudaGetLastError());
    } else if (g->n_padded > 0) {
        window_sort_kernel<<<(unsigned)(g->n_padded / WINDOW), 256, 0, s>>>(g->deg, g->pos, g->cnbc, g->maxnb, g->n_interior);
        count_launch();
        B200GEO_CUDA(cudaGetLastError());
    }
    // link slots per chunk -> first link slot of every chunk (the scan runs in place)
    chunk_width_kernel<<<blocks_for(chunks + 1), 256, 0, s>>>(g->cnbc, g->chunk_off, chunks);
    count_launch();
    B200GEO_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, need, g->chunk_off, g->chunk_off, (int)(chunks + 1), s));
    if ((rc = ensure_scan_tmp(g, need))) return rc;
    B200GEO_CUDA(cub::DeviceScan::ExclusiveSum(g->scan_tmp, need, g->chunk_off, g->chunk_off, (int)(chunks + 1), s));
    count_launch();
    int64_t link_slots = 0;
    unsigned long long links = 0;
    B200GEO_CUDA(cudaMemcpyAsync(&link_slots, g->chunk_off + chunks, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    B200GEO_CUDA(cudaMemcpyAsync(&links, g->link_count, sizeof(links), cudaMemcpyDeviceToHost, s));
    B200GEO_CUDA(cudaMemcpyAsync(err, g->err, sizeof(err), cudaMemcpyDeviceToHost, s));
    B200GEO_CUDA(cudaStreamSynchronize(s));
    if (err[2]) return fail(B200GEO_ERR_OUT_OF_RANGE, "ContainerCell capacity exeeded");
    if (err[3]) return fail(B200GEO_ERR_INVALID, "the ids of a container must ascend (ContainerCell::insert keeps them sorted)");
    g->n_link_slots = link_slots;
    g->n_links = (int64_t)links;
    // link slots beyond an item's own neighbour count (its chunk is as wide as the chunk's longest item) are never followed
    size_t link_bytes = (size_t)(link_slots > 0 ? link_slots : 1) * (g->kernel == 1 ? sizeof(uint16_t) : sizeof(int32_t));
    e = g->kernel == 1 ? cudaMalloc((void **)&g->link16, link_bytes) : cudaMalloc((void **)&g->link, link_bytes);
    if (e != cudaSuccess) {
        free_compact(g);
        cudaGetLastError();
        return fail(B200GEO_ERR_NOMEM, std::string("cudaMalloc failed: ") + cudaGetErrorString(e));
    }
    B200GEO_CUDA(cudaMemsetAsync(g->kernel == 1 ? (void *)g->link16 : (void *)g->link, 0, link_bytes, s));
    Dims dims;
    for (int i = 0; i < 3; ++i) {
        dims.d[i] = g->d[i];
        dims.wrap[i] = g->desc.ghost_mode[i][0] == B200GEO_GHOST_WRAP;
    }
    dims.ndims = g->ndims;
    int64_t slots = g->ncont * g->cap;
    if (g->kernel == 1) {
        B200GEO_CUDA(cudaMemsetAsync(g->cinflux, 0, np * sizeof(double), s));
        tile_fill_kernel<<<blocks_for(slots_all), 256, 0, s>>>(g->counts, g->offsets, g->pos, g->values, g->influx, g->cval[0], g->cval[1],
                                                               g->cinflux, g->cap, g->ncont);
        count_launch();
        B200GEO_CUDA(cudaGetLastError());
        tile_resolve_kernel<<<blocks_for(slots, 128), 128, 0, s>>>(g->counts, g->offsets, g->pos, g->ids, g->nbc, g->nbids, g->chunk_off,
                                                                   g->link16, g->maxnb, g->ncont, tiles, dims, g->err);
        if (g->tile_smem > 48 * 1024) {
            const int bytes = (int)g->tile_smem;
            B200GEO_CUDA(cudaFuncSetAttribute(tile_sweep_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
            B200GEO_CUDA(cudaFuncSetAttribute(tile_sweep_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
            B200GEO_CUDA(cudaFuncSetAttribute(tile_sweep_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
            B200GEO_CUDA(cudaFuncSetAttribute(tile_sweep_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        }
    } else {
        compact_kernel<<<blocks_for(slots_all), 256, 0, s>>>(g->counts, g->offsets, g->pos, g->values, g->influx, g->cval[0], g->cval[1],
                                                             g->cinflux, g->cap, g->ncont, g->n_interior, g->n_padded);
        count_launch();
        B200GEO_CUDA(cudaGetLastError());
        resolve_kernel<<<blocks_for(slots, 128), 128, 0, s>>>(g->counts, g->offsets, g->pos, g->ids, g->nbc, g->nbids, g->chunk_off, g->link,
                                                              g->cap, g->maxnb, g->ncont, g->n_interior, g->n_padded, dims, g->err);
    }
    count_launch();
    B200GEO_CUDA(cudaGetLastError());
    B200GEO_CUDA(cudaMemcpyAsync(err, g->err, sizeof(err), cudaMemcpyDeviceToHost, s));
    B200GEO_CUDA(cudaStreamSynchronize(s));
    if (err[0]) return fail(B200GEO_ERR_LOGIC, "id not found: could not find id " + std::to_string(err[1]) + " in neighborhood");
    g->cur = 0;
    g->dirty = false;
    g->slot_values_stale = false;
    ++g->rebuilds;
    return B200GEO_OK;
}

// n sweeps on the compact store
int sweeps(b200geo_containergrid *g, uint32_t n, cudaStream_t s)
{
    if (g->n_interior <= 0) return B200GEO_OK;
    if (g->kernel == 1) {
        Tiles tiles;
        tiles.g = g->tile_g;
        tiles.w = g->tile_w;
        tiles.cap = g->cap;
        tiles.rows = g->ndims == 3 ? 9 : 3;
        tiles.ndims = g->ndims;
        tiles.per_row = g->tiles_per_row;
        Dims dims;
        for (int i = 0; i < 3; ++i) {
            dims.d[i] = g->d[i];
            dims.wrap[i] = g->desc.ghost_mode[i][0] == B200GEO_GHOST_WRAP;
        }
        dims.ndims = g->ndims;
        const unsigned nt = (unsigned)g->n_tiles;
        for (uint32_t t = 0; t < n; ++t) {
            const double *from = g->cval[g->cur];
            double *to = g->cval[g->cur ^ 1];
#define B200GEO_TILE_SWEEP(CPW)                                                                                                  \
    tile_sweep_kernel<CPW><<<nt, 256, g->tile_smem, s>>>(from, to, g->cinflux, g->cnbc, g->ptarget, g->chunk_off, g->link16, g->offsets, \
                                                         g->ncont, tiles, dims)
            switch (g->tile_w / 256) {
            case 1: B200GEO_TILE_SWEEP(1); break;
            case 2: B200GEO_TILE_SWEEP(2); break;
            case 3: B200GEO_TILE_SWEEP(3); break;
            default: B200GEO_TILE_SWEEP(4); break;
            }
#undef B200GEO_TILE_SWEEP
            g->cur ^= 1;
        }
    } else {
        for (uint32_t t = 0; t < n; ++t) {
            sweep_kernel<<<blocks_for(g->n_interior), 256, 0, s>>>(g->cval[g->cur], g->cval[g->cur ^ 1], g->cinflux, g->cnbc, g->chunk_off,
                                                                   g->link, g->n_interior);
            g->cur ^= 1;
        }
    }
    count_launch(n);
    B200GEO_CUDA(cudaGetLastError());
    g->slot_values_stale = true;
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
    if (e == cudaSuccess) e = cudaMalloc((void **)&g->link_count, sizeof(unsigned long long));
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
    if (g->link_count) cudaFree(g->link_count);
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
    int rc = sweeps(g, n_steps, s);
    if (rc) return rc;
    g->sweeps += n_steps;
    return B200GEO_OK;
}

int b200geo_containergrid_stats(const b200geo_containergrid *g, uint64_t out[6])
{
    if (!g || !out) return fail(B200GEO_ERR_INVALID, "null argument");
    out[0] = (uint64_t)g->n_interior;
    out[1] = (uint64_t)g->n_links;
    out[2] = g->rebuilds;
    out[3] = g->sweeps;
    out[4] = (uint64_t)g->kernel;
    out[5] = (uint64_t)g->n_link_slots * (g->kernel == 1 ? sizeof(uint16_t) : sizeof(int32_t));
    return B200GEO_OK;
}

}
