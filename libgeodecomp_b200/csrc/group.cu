// Slab groups: the slabs of ONE simulation space on several GPUs of one box, driven by one host
// thread (b200geo_group_* in include/b200geo.h). This is the in-process twin of the one-rank-per-GPU
// NCCL path (libgeodecomp_b200/striping.py) and what the C++ façade's B200StripingSimulator calls.
//
// Replaces, for slab partitions along the last axis (geometry/partitions/stripingpartition.h:57-62):
//   StripingSimulator::nanoStep         parallelization/stripingsimulator.h:269-286 — update the rims,
//                                        start shipping them, update the interior while they travel
//   PatchLink::Accepter / Provider      communication/patchlink.h:127-151, 218-244 — MPI_Isend/Irecv of
//                                        saveRegion buffers -> direct device-to-device copies over
//                                        NVLink (cudaMemcpyPeerAsync) of the contiguous rim planes into
//                                        the neighbour's ghost planes: no pack kernel, no host staging
//   VanillaStepper ghost zone width k   parallelization/nesting/vanillastepper.h:157-225 — one exchange
//                                        per k sweeps; with the temporal-blocked Jacobi kernels the k
//                                        sweeps of a round are ONE launch per rim / interior
#include "grid.h"

#include <cstring>
#include <new>

#define B200GEO_GROUP_MAX 16

struct b200geo_group {
    int n;
    bool periodic;
    b200geo_grid *g[B200GEO_GROUP_MAX];
    cudaStream_t compute[B200GEO_GROUP_MAX], copy[B200GEO_GROUP_MAX];
    cudaEvent_t rim[B200GEO_GROUP_MAX], copied[B200GEO_GROUP_MAX];
    int valid;  // ghost planes valid on every PEER side of the current buffers
    uint64_t exchanges, bytes_moved;
};

namespace b200geo {

namespace {

// What updates a box of one slab: a kernel family of this library (possibly several fused sweeps per
// launch), or a callback for cells whose kernel lives in the caller's translation unit (one sweep).
struct Updater {
    int kernel;
    const void *params;
    b200geo_update_fn fn;
    void *ctx;

    int box(b200geo_grid *g, uint32_t nano_step, const int32_t origin[3], const int32_t dim[3], int sweeps, cudaStream_t s) const
    {
        if (!fn) return b200geo_update_box_n(g, kernel, params, nano_step, origin, dim, (uint32_t)sweeps, s);
        if (sweeps != 1) return fail(B200GEO_ERR_LOGIC, "update callbacks take one sweep per call");
        int rc = fn(ctx, g, nano_step, origin, dim, s);
        if (rc < 0) return fail(rc, "update callback failed");
        return B200GEO_OK;
    }
};

// neighbour of slab i on `side` (0 = low, 1 = high), or -1
int neighbour(const b200geo_group *grp, int i, int side)
{
    int j = side == 0 ? i - 1 : i + 1;
    if (j < 0) j = grp->periodic ? grp->n - 1 : -1;
    if (j >= grp->n) j = grp->periodic ? 0 : -1;
    if (grp->n == 1) j = -1;
    return j;
}

enum { LBM_T = 5, LBM_B = 6, LBM_SE = 10, LBM_TW = 11, LBM_BW = 12, LBM_TE = 13, LBM_BE = 14, LBM_TN = 15, LBM_BN = 16, LBM_TS = 17, LBM_BS = 18 };

// The members the update reads from the ghost planes on `side`. LBM with ghost width 1 reads only the
// populations that cross the face (csrc/lbm.cu: T* pulled from z - 1; B* from z + 1, plus SE which the
// EAST_NOSLIP rule takes from z + 1) — everything else of a ghost cell is never read between two
// exchanges. Wider ghost zones recompute the rim and need whole cells.
int needed_members(const b200geo_grid *g, int kernel, int width, int side, int *out)
{
    if (kernel == B200GEO_KERNEL_LBM_D3Q19 && width == 1 && g->n == 24) {
        static const int low[] = {LBM_T, LBM_TW, LBM_TE, LBM_TN, LBM_TS};
        static const int high[] = {LBM_B, LBM_BW, LBM_BE, LBM_BN, LBM_BS, LBM_SE};
        int n = side == 0 ? 5 : 6;
        memcpy(out, side == 0 ? low : high, n * sizeof(int));
        return n;
    }
    if (kernel == B200GEO_KERNEL_LBM_D3Q19 && g->n == 24) {
        // the rim is recomputed: 19 populations and the state; density / velocity of a ghost cell are never read
        for (int m = 0; m < 19; ++m) out[m] = m;
        out[19] = 23;
        return 20;
    }
    for (int m = 0; m < g->n; ++m) out[m] = m;
    return g->n;
}

// Ship the `width` outermost owned planes of every slab (buffer `which`: 0 = current, 1 = scratch) into
// the matching ghost planes of its neighbours (same buffer index over there: the slabs step in lock
// step). Enqueued on the copy streams; the caller has made them wait for the producers.
int ship_rims(b200geo_group *grp, int kernel, int width, int which)
{
    for (int i = 0; i < grp->n; ++i) {
        b200geo_grid *g = grp->g[i];
        B200GEO_CUDA(cudaSetDevice(g->device));
        for (int side = 0; side < 2; ++side) {
            int j = neighbour(grp, i, side);
            if (j < 0) continue;
            b200geo_grid *peer = grp->g[j];
            // our low-side rim is the neighbour's HIGH ghost, so it wants what its high side reads
            int members[B200GEO_MAX_MEMBERS];
            int n = needed_members(g, kernel, width, 1 - side, members);
            for (int k = 0; k < n; ++k) {
                void *src = 0, *dst = 0;
                uint64_t bytes = 0, dst_bytes = 0;
                int rc = b200geo_halo_block_in(g, members[k], side, 0, width, which, &src, &bytes);
                if (rc) return rc;
                rc = b200geo_halo_block_in(peer, members[k], 1 - side, 1, width, which, &dst, &dst_bytes);
                if (rc) return rc;
                if (bytes != dst_bytes) return fail(B200GEO_ERR_INVALID, "slabs of one group must have the same cross-section");
                B200GEO_CUDA(cudaMemcpyPeerAsync(dst, peer->device, src, g->device, bytes, grp->copy[i]));
                grp->bytes_moved += bytes;
            }
        }
        B200GEO_CUDA(cudaEventRecord(grp->copied[i], grp->copy[i]));
    }
    ++grp->exchanges;
    return B200GEO_OK;
}

// every compute stream waits for the copies that read its rims or wrote its ghosts
int wait_for_copies(b200geo_group *grp)
{
    for (int i = 0; i < grp->n; ++i) {
        B200GEO_CUDA(cudaSetDevice(grp->g[i]->device));
        B200GEO_CUDA(cudaStreamWaitEvent(grp->compute[i], grp->copied[i], 0));
        for (int side = 0; side < 2; ++side) {
            int j = neighbour(grp, i, side);
            if (j >= 0) B200GEO_CUDA(cudaStreamWaitEvent(grp->compute[i], grp->copied[j], 0));
        }
    }
    return B200GEO_OK;
}

int mark_valid(b200geo_group *grp, int width)
{
    for (int i = 0; i < grp->n; ++i)
        for (int side = 0; side < 2; ++side)
            if (neighbour(grp, i, side) >= 0) grp->g[i]->peer_valid[side] = width;
    grp->valid = width;
    return B200GEO_OK;
}

// sweeps a kernel family can take per launch (b200geo_update_box_n): jacobi_tb.cu, lbm_tb.cu
int max_fused_sweeps(int kernel)
{
    if (kernel == B200GEO_KERNEL_JACOBI6 || kernel == B200GEO_KERNEL_JACOBI7 || kernel == B200GEO_KERNEL_JACOBI27) return 4;
    return kernel == B200GEO_KERNEL_LBM_D3Q19 ? 2 : 1;
}

int ghost_width(const b200geo_group *grp)
{
    const b200geo_grid *g = grp->g[0];
    return g->g[g->slab_axis];
}

// blocking exchange of the current buffers (all members): rims -> neighbours' ghosts
int exchange_current(b200geo_group *grp, int kernel, int width)
{
    for (int i = 0; i < grp->n; ++i) {
        B200GEO_CUDA(cudaSetDevice(grp->g[i]->device));
        B200GEO_CUDA(cudaEventRecord(grp->rim[i], grp->compute[i]));
    }
    for (int i = 0; i < grp->n; ++i) {
        B200GEO_CUDA(cudaSetDevice(grp->g[i]->device));
        B200GEO_CUDA(cudaStreamWaitEvent(grp->copy[i], grp->rim[i], 0));
        for (int side = 0; side < 2; ++side) {
            int j = neighbour(grp, i, side);
            if (j >= 0) B200GEO_CUDA(cudaStreamWaitEvent(grp->copy[i], grp->rim[j], 0));
        }
    }
    int rc = ship_rims(grp, kernel, width, 0);
    if (rc) return rc;
    rc = wait_for_copies(grp);
    if (rc) return rc;
    return mark_valid(grp, width);
}

// One round of `sweeps` sweeps, rim first: the `w` outermost planes on both sides of every slab, ship them out
// of the scratch buffers while the interiors are updated, swap. Bound kernels fuse sweeps = w sweeps per launch;
// callbacks take one sweep per round (w = their stencil radius).
int overlapped_round(b200geo_group *grp, const Updater& up, uint32_t nano_step, int w, int sweeps)
{
    int a = grp->g[0]->slab_axis;
    for (int i = 0; i < grp->n; ++i) {
        b200geo_grid *g = grp->g[i];
        B200GEO_CUDA(cudaSetDevice(g->device));
        int rc = refresh_wrap(g, grp->compute[i]);
        if (rc) return rc;
        int n = g->d[a];
        int lo = neighbour(grp, i, 0) >= 0 ? w : 0, hi = neighbour(grp, i, 1) >= 0 ? n - w : n;
        int32_t origin[3] = {0, 0, 0}, dim[3] = {g->d[0], g->d[1], g->d[2]};
        if (lo > 0) {
            origin[a] = 0;
            dim[a] = lo;
            rc = up.box(g, nano_step, origin, dim, sweeps, grp->compute[i]);
            if (rc) return rc;
        }
        if (hi < n) {
            origin[a] = hi;
            dim[a] = n - hi;
            rc = up.box(g, nano_step, origin, dim, sweeps, grp->compute[i]);
            if (rc) return rc;
        }
        B200GEO_CUDA(cudaEventRecord(grp->rim[i], grp->compute[i]));
    }
    // a copy may start once its source rim is written AND the destination slab has left the previous
    // round (its rim event of this round is behind all of that round's kernels on its stream)
    for (int i = 0; i < grp->n; ++i) {
        B200GEO_CUDA(cudaSetDevice(grp->g[i]->device));
        B200GEO_CUDA(cudaStreamWaitEvent(grp->copy[i], grp->rim[i], 0));
        for (int side = 0; side < 2; ++side) {
            int j = neighbour(grp, i, side);
            if (j >= 0) B200GEO_CUDA(cudaStreamWaitEvent(grp->copy[i], grp->rim[j], 0));
        }
    }
    int rc = ship_rims(grp, up.kernel, w, 1);
    if (rc) return rc;
    for (int i = 0; i < grp->n; ++i) {
        b200geo_grid *g = grp->g[i];
        B200GEO_CUDA(cudaSetDevice(g->device));
        int n = g->d[a];
        int lo = neighbour(grp, i, 0) >= 0 ? w : 0, hi = neighbour(grp, i, 1) >= 0 ? n - w : n;
        if (hi > lo) {
            int32_t origin[3] = {0, 0, 0}, dim[3] = {g->d[0], g->d[1], g->d[2]};
            origin[a] = lo;
            dim[a] = hi - lo;
            rc = up.box(g, nano_step, origin, dim, sweeps, grp->compute[i]);
            if (rc) return rc;
        }
        g->cur ^= 1;
        g->sweeps += sweeps;
    }
    rc = wait_for_copies(grp);
    if (rc) return rc;
    return mark_valid(grp, w);
}

// callbacks on slabs too thin for a rim / interior split: exchange, one sweep over every slab, swap
int plain_round(b200geo_group *grp, const Updater& up, uint32_t nano_step, int w)
{
    int rc = exchange_current(grp, 0, w);
    if (rc) return rc;
    for (int i = 0; i < grp->n; ++i) {
        b200geo_grid *g = grp->g[i];
        B200GEO_CUDA(cudaSetDevice(g->device));
        rc = refresh_wrap(g, grp->compute[i]);
        if (rc) return rc;
        int32_t origin[3] = {0, 0, 0}, dim[3] = {g->d[0], g->d[1], g->d[2]};
        rc = up.box(g, nano_step, origin, dim, 1, grp->compute[i]);
        if (rc) return rc;
        g->cur ^= 1;
        g->sweeps += 1;
    }
    for (int i = 0; i < grp->n; ++i) grp->g[i]->peer_valid[0] = grp->g[i]->peer_valid[1] = 0;
    grp->valid = 0;
    return B200GEO_OK;
}

}

}

using namespace b200geo;

extern "C" {

int b200geo_group_create(b200geo_grid *const *grids, int n, int periodic, b200geo_group **out)
{
    if (!grids || !out || n < 1 || n > B200GEO_GROUP_MAX) return fail(B200GEO_ERR_INVALID, "bad slab group");
    for (int i = 0; i < n; ++i) {
        const b200geo_grid *g = grids[i];
        if (!g) return fail(B200GEO_ERR_INVALID, "null grid in slab group");
        int a = g->slab_axis;
        if (g->slab_axis != grids[0]->slab_axis || g->n != grids[0]->n || g->g[a] != grids[0]->g[grids[0]->slab_axis])
            return fail(B200GEO_ERR_INVALID, "slabs of one group must agree in layout and ghost width");
        for (int k = 0; k < 3; ++k)
            if (k != a && (g->d[k] != grids[0]->d[k] || g->g[k] != grids[0]->g[k]))
                return fail(B200GEO_ERR_INVALID, "slabs of one group must have the same cross-section");
        if (n > 1) {
            bool low = i > 0 || periodic, high = i < n - 1 || periodic;
            if ((g->desc.ghost_mode[a][0] == B200GEO_GHOST_PEER) != low || (g->desc.ghost_mode[a][1] == B200GEO_GHOST_PEER) != high)
                return fail(B200GEO_ERR_INVALID, "slab faces towards a neighbour must be PEER ghost layers, outer faces must not");
            if (g->d[a] < g->g[a]) return fail(B200GEO_ERR_INVALID, "slab thinner than the ghost zone");
        }
    }
    b200geo_group *grp = new (std::nothrow) b200geo_group();
    if (!grp) return fail(B200GEO_ERR_NOMEM, "out of host memory");
    memset(grp, 0, sizeof(*grp));
    grp->n = n;
    grp->periodic = periodic != 0;
    for (int i = 0; i < n; ++i) grp->g[i] = grids[i];
    for (int i = 0; i < n; ++i) {
        int dev = grids[i]->device;
        cudaError_t e = cudaSetDevice(dev);
        // direct NVLink access to the slab neighbours (without it cudaMemcpyPeerAsync stages through the host)
        for (int side = 0; side < 2 && e == cudaSuccess; ++side) {
            int j = neighbour(grp, i, side);
            if (j < 0 || grids[j]->device == dev) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, dev, grids[j]->device) == cudaSuccess && can) {
                cudaError_t p = cudaDeviceEnablePeerAccess(grids[j]->device, 0);
                if (p != cudaSuccess && p != cudaErrorPeerAccessAlreadyEnabled) e = p;
                cudaGetLastError();
            }
        }
        // blocking streams: ordered against the legacy default stream the grid I/O calls use
        if (e == cudaSuccess) e = cudaStreamCreate(&grp->compute[i]);
        if (e == cudaSuccess) e = cudaStreamCreate(&grp->copy[i]);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&grp->rim[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&grp->copied[i], cudaEventDisableTiming);
        if (e != cudaSuccess) {
            b200geo_group_destroy(grp);
            return check_cuda(e, "slab group setup");
        }
    }
    *out = grp;
    return B200GEO_OK;
}

int b200geo_group_destroy(b200geo_group *grp)
{
    if (!grp) return B200GEO_OK;
    for (int i = 0; i < grp->n; ++i) {
        cudaSetDevice(grp->g[i]->device);
        if (grp->compute[i]) { cudaStreamSynchronize(grp->compute[i]); cudaStreamDestroy(grp->compute[i]); }
        if (grp->copy[i]) { cudaStreamSynchronize(grp->copy[i]); cudaStreamDestroy(grp->copy[i]); }
        if (grp->rim[i]) cudaEventDestroy(grp->rim[i]);
        if (grp->copied[i]) cudaEventDestroy(grp->copied[i]);
    }
    delete grp;
    return B200GEO_OK;
}

int b200geo_group_invalidate(b200geo_group *grp)
{
    if (!grp) return fail(B200GEO_ERR_INVALID, "null group");
    for (int i = 0; i < grp->n; ++i) grp->g[i]->peer_valid[0] = grp->g[i]->peer_valid[1] = 0;
    grp->valid = 0;
    return B200GEO_OK;
}

int b200geo_group_exchange(b200geo_group *grp)
{
    if (!grp) return fail(B200GEO_ERR_INVALID, "null group");
    if (grp->n == 1) return B200GEO_OK;
    return exchange_current(grp, 0, ghost_width(grp));
}

int b200geo_group_step(b200geo_group *grp, int kernel, const void *params, uint32_t first_nano_step, uint32_t n_steps)
{
    if (!grp) return fail(B200GEO_ERR_INVALID, "null group");
    const int w = ghost_width(grp), a = grp->g[0]->slab_axis;
    // LBM, default parameter block: density / velocity are stored by the last sweep of this call only
    const bool lbm_lazy = kernel == B200GEO_KERNEL_LBM_D3Q19 && (!params || *(const int32_t *)params == 0);
    const int32_t lbm_store = 0, lbm_skip = 2;
    bool overlap = w == 1 || w <= max_fused_sweeps(kernel);
    for (int i = 0; i < grp->n; ++i)
        if (grp->g[i]->d[a] < 2 * w) overlap = false;
    uint32_t done = 0;
    while (done < n_steps) {
        uint32_t left = n_steps - done;
        if (grp->n == 1) {
            B200GEO_CUDA(cudaSetDevice(grp->g[0]->device));
            return b200geo_step(grp->g[0], kernel, params, first_nano_step + done, left, grp->compute[0]);
        }
        if (grp->valid == 0 || (overlap && grp->valid < w)) {
            // first exchange (or after an invalidate): whole cells, so that members the update never
            // re-reads from a neighbour (wall states, ...) are in place as well. Leftover validity from an
            // earlier call (0 < valid < w) is topped up the same way, so that every round below is a fused,
            // overlapped one (StripingSimulator::nanoStep, parallelization/stripingsimulator.h:269-286)
            int rc = exchange_current(grp, 0, w);
            if (rc) return rc;
        }
        if (overlap) {
            // a round of fewer than w sweeps (the tail of this call) still ships w planes: the ghost zones are
            // w deep again afterwards and the next call starts with an overlapped round, not with an exchange
            const uint32_t k = left < (uint32_t)w ? left : (uint32_t)w;
            const void *p = lbm_lazy ? (const void *)(done + k == n_steps ? &lbm_store : &lbm_skip) : params;
            Updater up = {kernel, p, 0, 0};
            int rc = overlapped_round(grp, up, first_nano_step + done, w, (int)k);
            if (rc) return rc;
            done += k;
            continue;
        }
        uint32_t k = left < (uint32_t)grp->valid ? left : (uint32_t)grp->valid;
        const void *p = lbm_lazy ? (const void *)(done + k == n_steps ? &lbm_store : &lbm_skip) : params;
        for (int i = 0; i < grp->n; ++i) {
            B200GEO_CUDA(cudaSetDevice(grp->g[i]->device));
            int rc = b200geo_step(grp->g[i], kernel, p, first_nano_step + done, k, grp->compute[i]);
            if (rc) return rc;
        }
        grp->valid -= (int)k;
        done += k;
    }
    return B200GEO_OK;
}

int b200geo_group_step_with(b200geo_group *grp, b200geo_update_fn update, void *ctx, uint32_t first_nano_step, uint32_t n_steps)
{
    if (!grp || !update) return fail(B200GEO_ERR_INVALID, "null argument");
    const int w = ghost_width(grp), a = grp->g[0]->slab_axis;
    Updater up = {0, 0, update, ctx};
    bool overlap = grp->n > 1;
    for (int i = 0; i < grp->n; ++i)
        if (grp->g[i]->d[a] < 2 * w) overlap = false;
    for (uint32_t done = 0; done < n_steps; ++done) {
        const uint32_t nano = first_nano_step + done;
        if (grp->n == 1) {
            b200geo_grid *g = grp->g[0];
            B200GEO_CUDA(cudaSetDevice(g->device));
            int rc = refresh_wrap(g, grp->compute[0]);
            if (rc) return rc;
            int32_t origin[3] = {0, 0, 0}, dim[3] = {g->d[0], g->d[1], g->d[2]};
            rc = up.box(g, nano, origin, dim, 1, grp->compute[0]);
            if (rc) return rc;
            g->cur ^= 1;
            g->sweeps += 1;
            continue;
        }
        if (!overlap) {
            int rc = plain_round(grp, up, nano, w);
            if (rc) return rc;
            continue;
        }
        if (grp->valid < w) {
            int rc = exchange_current(grp, 0, w);
            if (rc) return rc;
        }
        int rc = overlapped_round(grp, up, nano, w, 1);
        if (rc) return rc;
    }
    return B200GEO_OK;
}

int b200geo_group_sync(b200geo_group *grp)
{
    if (!grp) return fail(B200GEO_ERR_INVALID, "null group");
    for (int i = 0; i < grp->n; ++i) {
        B200GEO_CUDA(cudaSetDevice(grp->g[i]->device));
        B200GEO_CUDA(cudaStreamSynchronize(grp->copy[i]));
        B200GEO_CUDA(cudaStreamSynchronize(grp->compute[i]));
    }
    return B200GEO_OK;
}

int b200geo_group_stats(const b200geo_group *grp, uint64_t out[2])
{
    if (!grp || !out) return fail(B200GEO_ERR_INVALID, "null argument");
    out[0] = grp->exchanges;
    out[1] = grp->bytes_moved;
    return B200GEO_OK;
}

}
