/* b200geo — C ABI of the B200-native cell-update engine (libb200geo.so).
 *
 * LibGeoDecomp has no FFI boundary of its own: back-ends plug in as C++ templates
 * (Simulator<CELL> subclasses, parallelization/cudasimulator.h:298; Stepper<CELL> subclasses,
 * parallelization/nesting/cudastepper.h:256). This header is the seam a maintainer binds
 * instead: plain pointers, sizes and status codes, called only by the header-only façade
 * templates in include/libgeodecomp_b200/ (B200Simulator<CELL>, B200Grid<CELL>) and by the
 * ctypes mirror in libgeodecomp_b200/capi.py. Each entry point names the reference interface it
 * replaces. Paths are relative to /root/reference/src/libgeodecomp/.
 *
 * Conventions
 *  - every function returns B200GEO_OK (0) or a negative b200geo_status; the façade maps them to
 *    the reference's exceptions (std::invalid_argument, std::logic_error, std::out_of_range,
 *    std::runtime_error("CUDA error"), misc/cudautil.h:48-55). b200geo_last_error() has the text.
 *  - coordinates are interior cell coordinates, x fastest; 2-D grids use dim[2] = 1.
 *  - a streak is int32 {x, y, z, endX} (geometry/streak.h:16-84, half open in x).
 *  - "member-major" buffers are byte-compatible with SoAGrid::saveRegion / loadRegion
 *    (storage/soagrid.h:523-576, storage/serializationbuffer.h:61-101): for each member in
 *    registration order all cells of the region, tightly packed.
 *  - a `stream` argument is a cudaStream_t passed as void* (NULL = legacy default stream).
 *  - not re-entrant per handle; one host thread drives one device. There is NO CPU fallback:
 *    without a usable CUDA device every compute entry point returns B200GEO_ERR_CUDA.
 */
#ifndef B200GEO_H
#define B200GEO_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200GEO_MAX_MEMBERS 32

typedef enum {
    B200GEO_OK = 0,
    B200GEO_ERR_INVALID = -1,      /* -> std::invalid_argument */
    B200GEO_ERR_LOGIC = -2,        /* -> std::logic_error (unsupported combination) */
    B200GEO_ERR_OUT_OF_RANGE = -3, /* -> std::out_of_range (e.g. particle capacity exceeded, storage/fixedarray.h:77-83) */
    B200GEO_ERR_CUDA = -4,         /* -> std::runtime_error("CUDA error") */
    B200GEO_ERR_NOMEM = -5         /* -> std::bad_alloc */
} b200geo_status;

/* MemoryLocation::Location, storage/memorylocation.h:11-14 */
typedef enum { B200GEO_HOST = 0, B200GEO_CUDA_DEVICE = 1 } b200geo_location;

/* Hand-written kernel families; B200KernelBinding<CELL> maps a user cell type to one of these. */
typedef enum {
    B200GEO_KERNEL_JACOBI6 = 1,  /* 6-point mean, f64; src/examples/jacobi3d/main.cpp:30-39 */
    B200GEO_KERNEL_JACOBI7 = 2,  /* 7-point mean, f64; src/testbed/performancetests/main.cpp:1292-1299 */
    B200GEO_KERNEL_JACOBI27 = 3, /* 27-point mean over Moore<3,1>, f64; oracle/models/jacobi.h */
    B200GEO_KERNEL_GOL = 4,      /* Conway's Life, 1-byte cells; src/examples/gameoflife/main.cpp:36-59 */
    B200GEO_KERNEL_LBM_D3Q19 = 5,/* D3Q19 BGK + wall states, f32; src/examples/latticeboltzmann/main.cpp:62-229 */
    B200GEO_KERNEL_NBODY = 6,    /* BoxCell short-range n-body; storage/boxcell.h:112-174 */
    B200GEO_KERNEL_CONTAINER = 7 /* ContainerCell ID-keyed cargo (meshfree / unstructured); storage/containercell.h:170-200 */
} b200geo_kernel;

/* What a ghost (padding) layer on one side of one axis holds. */
typedef enum {
    B200GEO_GHOST_EDGE = 0, /* the constant edge cell of a Cube axis (storage/soagrid.h:578-584) */
    B200GEO_GHOST_WRAP = 1, /* periodic image of this grid's own far side (Torus axis on one device) */
    B200GEO_GHOST_PEER = 2  /* cells owned by a neighbouring subdomain, filled by the halo exchange */
} b200geo_ghost_mode;

/* Describes one device-resident double-buffered SoA grid. Replaces the ctor arguments of
 * SoAGrid (storage/soagrid.h:344-427) / CUDASoAGrid (storage/cudasoagrid.h:116) plus the SoA
 * member table LIBFLATARRAY_REGISTER_SOA generates (lib/libflatarray/.../macros.hpp:101-126). */
typedef struct {
    int32_t dim[3];                            /* interior extent */
    int32_t ghost[3];                          /* ghost width per axis (>= stencil radius where used, 0 otherwise) */
    int32_t ghost_mode[3][2];                  /* b200geo_ghost_mode per axis and side (0 = low, 1 = high) */
    int32_t n_members;
    int32_t member_bytes[B200GEO_MAX_MEMBERS]; /* 1, 2, 4 or 8; array members are listed element by element */
} b200geo_grid_desc;

typedef struct b200geo_grid b200geo_grid;

/* ---- library ---------------------------------------------------------------------------- */
const char *b200geo_version(void);
const char *b200geo_last_error(void);
/* number of CUDA devices, or a negative status (no driver / no device). */
int b200geo_device_count(void);
/* Launch / tiling parameters by name (the role misc/cudasimulationfactory.h:28-33 BlockDimX/Y/Z
 * play for the reference's CUDASimulator). Unknown keys -> B200GEO_ERR_INVALID. Keys:
 * "jacobi.zchunk", "jacobi.prefetch", "gol.rows", "lbm.block", "jacobi.tb" (sweeps fused per launch
 * by the temporal-blocked Jacobi kernels, 1..4; 0 = automatic: 2 for the 27-point kernel, 4 for 6/7-point), "jacobi.tb_rows" (tile shape), "jacobi.tb_zchunk", "gol.bits" (fewest sweeps per call that run
 * bit-packed; 0 = never), "gol.bits_rows", "nbody.kernel", "jacobi.pdl" (programmatic dependent launch of the one-sweep Jacobi
 * kernel: 0, 1, < 0 = small grids only), "lbm.variant" (rows per thread of the LBM kernel: 1 or 2), "jacobi.tb_promo" (L2 promotion of the
 * temporal-blocked kernel's TMA loads), "jacobi.tb_raster" (its CTA order), "nbody.run" (containers per CTA), "jacobi.resident" (1: small Cube grids take the
 * SM-resident multi-sweep kernel), "container.kernel" (ContainerCell sweeps: 0 = links into the global value array, 1 = tiles of
 * containers with the neighbourhood's values staged in shared memory and 16-bit links; takes effect at the next link resolution).
 * value < 0 restores the default. */
int b200geo_set_tuning(const char *key, int value);
/* number of kernels this library has launched so far in this process (bench.py: gpu_launches). */
uint64_t b200geo_launch_count(void);

/* ---- grid life cycle: replaces SoAGrid::SoAGrid/resize (storage/soagrid.h:380-427) -------- */
int b200geo_grid_create(const b200geo_grid_desc *desc, int device, b200geo_grid **out);
/* Same grid in the UNIFORM ELEMENT LAYOUT: all members share one lead-in and one row / plane pitch counted in
 * elements, and every member array holds exactly `member_stride` elements, so member m starts
 * member_stride x (bytes of the members before it) into a buffer. That is the addressing contract of
 * LibFlatArray's soa_accessor (lib/libflatarray/include/libflatarray/macros.hpp:327-349: data + DIM_PROD x
 * offset<CELL, m> + index x sizeof(member)) with DIM_PROD = member_stride — what lets the generic device path run a
 * model's SoA-signature updateLineX() with the accessors LIBFLATARRAY_REGISTER_SOA generated
 * (include/libgeodecomp_b200/b200genericsoa.h). member_stride: a multiple of 256, at least
 * b200geo_grid_uniform_min_stride(desc). b200geo_grid_member_ptr(g, 0, which) is the accessors' data pointer. */
int b200geo_grid_create_uniform(const b200geo_grid_desc *desc, int device, int64_t member_stride, b200geo_grid **out);
int b200geo_grid_uniform_min_stride(const b200geo_grid_desc *desc, int64_t *min_stride);
/* How b200geo_grid_create (member_stride = 0) / b200geo_grid_create_uniform would lay the member arrays out, without
 * allocating anything — needs no device. layout: int64 [n_members][7] = element bytes, lead-in (elements before
 * interior x = 0 in a row), row pitch, plane pitch, element offset of interior cell (0,0,0), bytes of the member array,
 * byte offset of the member array inside a buffer. Fails like the create calls for a bad description. */
int b200geo_grid_plan(const b200geo_grid_desc *desc, int64_t member_stride, int64_t *layout, int64_t *buffer_bytes);
/* elements per member array of a uniform-layout grid; 0 for a grid in the default layout */
int b200geo_grid_member_stride(const b200geo_grid *g, int64_t *member_stride);
int b200geo_grid_destroy(b200geo_grid *g);
/* bytes of one of the two buffers (padded layout) */
int b200geo_grid_buffer_bytes(const b200geo_grid *g, uint64_t *bytes);
/* the CUDA device the grid lives on */
int b200geo_grid_device(const b200geo_grid *g, int *device);
/* layout query for one member: row pitch and plane pitch in elements, and the element offset of
 * interior cell (0,0,0) from the member's base pointer */
int b200geo_grid_layout(const b200geo_grid *g, int member, int64_t *pitch_x, int64_t *pitch_plane,
                        int64_t *origin_offset);
/* device pointer of member `member` in the current (which = 0) or the scratch (which = 1) buffer */
int b200geo_grid_member_ptr(const b200geo_grid *g, int member, int which, void **ptr);

/* SoAGrid::setEdge/getEdge (storage/soagrid.h:486-499): cell = aggregated member bytes in
 * registration order; rewrites every EDGE ghost layer of BOTH buffers. */
int b200geo_grid_set_edge(b200geo_grid *g, const void *cell, void *stream);
int b200geo_grid_get_edge(const b200geo_grid *g, void *cell);

/* ---- bulk I/O --------------------------------------------------------------------------- */
/* GridBase::loadMember/saveMember (storage/gridbase.h:217-261) for a box: dense [dz][dy][dx]
 * array of one member <-> grid. `both` != 0 writes both buffers (SerialSimulator initialises
 * both grids, parallelization/serialsimulator.h:54-57). Host buffers may be pageable or pinned. */
int b200geo_grid_load_member(b200geo_grid *g, int member, const int32_t origin[3], const int32_t dim[3],
                             const void *src, int location, int both, void *stream);
int b200geo_grid_save_member(const b200geo_grid *g, int member, const int32_t origin[3], const int32_t dim[3],
                             void *dst, int location, void *stream);
/* GridBase::loadRegion/saveRegion(std::vector<char>) (storage/gridbase.h:32-50, soagrid.h:523-576,
 * LFA detail/save_functor.hpp:26-66 which launches one kernel PER STREAK on CUDA): here one
 * launch per call for any number of streaks. buf is member-major. */
int b200geo_grid_load_region(b200geo_grid *g, const int32_t *streaks, int n_streaks,
                             const void *buf, int location, int both, void *stream);
int b200geo_grid_save_region(const b200geo_grid *g, const int32_t *streaks, int n_streaks,
                             void *buf, int location, void *stream);

/* ---- the hot path ----------------------------------------------------------------------- */
/* n_steps x { UpdateFunctor over the whole grid; swap } = SerialSimulator::nanoStep
 * (parallelization/serialsimulator.h:132-139, storage/updatefunctor.h:403-428) on the device.
 * WRAP ghost layers are refreshed internally before every sweep; PEER layers must have been
 * filled by the caller (b200geo_halo_*): with ghost width G on a PEER side up to G steps may be
 * taken between two exchanges (the rim is recomputed redundantly, as
 * parallelization/nesting/vanillastepper.h:157-225 does). params: kernel specific, may be NULL. */
int b200geo_step(b200geo_grid *g, int kernel, const void *params, uint32_t first_nano_step,
                 uint32_t n_steps, void *stream);
/* Same, restricted to a box of interior coordinates (ghost cells addressable with negative
 * coordinates); does not swap. Used for rim-first / interior-overlapped schedules
 * (parallelization/stripingsimulator.h:269-286). */
int b200geo_update_box(b200geo_grid *g, int kernel, const void *params, uint32_t nano_step,
                       const int32_t origin[3], const int32_t dim[3], void *stream);
/* Same for n_sweeps fused sweeps of the box in ONE launch (temporal-blocked kernels: Jacobi, 2..4
 * sweeps; n_sweeps = 1 is b200geo_update_box). Ghost / neighbour cells must be valid n_sweeps deep
 * around the box. Reads the current buffer, writes the scratch buffer, does not swap: lets a slab
 * update its rims first, ship them, and overlap the interior with the transfer. Kernels that cannot
 * fuse sweeps -> B200GEO_ERR_LOGIC. */
int b200geo_update_box_n(b200geo_grid *g, int kernel, const void *params, uint32_t nano_step,
                         const int32_t origin[3], const int32_t dim[3], uint32_t n_sweeps, void *stream);
int b200geo_swap(b200geo_grid *g);
/* Re-materialise the periodic images of WRAP axes in the current buffer (done automatically by
 * b200geo_step; needed before b200geo_update_box on Torus axes). */
int b200geo_refresh_ghosts(b200geo_grid *g, void *stream);
int b200geo_sync(void *stream);
/* wait for `stream` of the device this grid lives on (b200geo_sync waits on the calling thread's CURRENT device: a host
 * thread that drives several GPUs — slab groups — uses this one) */
int b200geo_grid_sync(const b200geo_grid *g, void *stream);
/* Streams of a device, for callers that overlap host <-> device copies with sweeps (a streamed run: the Initializer
 * fills planes the sweeps have not reached while finished planes travel to the ParallelWriters — the reference hands
 * both plugins sub-boxes of the grid, io/initializer.h:38-71, io/parallelwriter.h:92-99). Every entry point that takes
 * a `stream` accepts these handles (they are cudaStream_t). b200geo_stream_wait: work enqueued on `waiter` from now
 * on starts after everything enqueued on `signaller` so far; the host does not wait. */
int b200geo_stream_create(int device, void **stream);
int b200geo_stream_destroy(int device, void *stream);
int b200geo_stream_wait(int device, void *waiter, void *signaller);
/* Plain device memory on `device` for region buffers that stay on the GPU: the device twins of the host-side
 * PatchBufferFixed a Stepper keeps for its rim and its volatile kernel (parallelization/nesting/commonstepper.h:
 * 29-30, 236-283; storage/patchbufferfixed.h) — filled and drained by b200geo_grid_save_region /
 * _load_region with location = B200GEO_CUDA_DEVICE. */
int b200geo_device_alloc(int device, uint64_t bytes, void **ptr);
int b200geo_device_free(int device, void *ptr);
/* Copy between such buffers: device to device across GPUs (over NVLink where the GPUs are peers — what a PatchLink
 * between two steppers of one process does instead of MPI_Isend / MPI_Irecv, communication/patchlink.h:127-151,
 * 218-244), or to / from host memory (device index < 0). Ordered after everything enqueued so far on the null streams
 * of both devices and before whatever is enqueued there afterwards. */
int b200geo_device_copy(int dst_device, void *dst, int src_device, const void *src, uint64_t bytes);
/* Page-locked host memory: Initializers / Writers that hand the engine whole boxes (GridBase::loadMember /
 * saveMember, storage/gridbase.h:217-261) out of such a buffer move them at the full speed of the link; pageable
 * memory works as well, only slower (the driver stages it). */
int b200geo_host_alloc(uint64_t bytes, void **ptr);
int b200geo_host_free(void *ptr);

/* ---- halo exchange (replaces PatchLink::Accepter::put / Provider::get,
 *      communication/patchlink.h:127-151,218-244, for slab partitions along the last axis,
 *      geometry/partitions/stripingpartition.h:57-62) --------------------------------------- */
/* Device address and byte count of the contiguous block of `width` whole padded planes of one
 * member: kind 0 = the outermost OWNED planes on `side` (what a neighbour needs = inner ghost
 * zone), kind 1 = the GHOST planes on `side` (outer ghost zone). Buffer = current. */
int b200geo_halo_block(const b200geo_grid *g, int member, int side, int kind, int width,
                       void **ptr, uint64_t *bytes);
/* Same in the current (which = 0) or the scratch (which = 1) buffer: a rim-first schedule ships
 * the freshly written rims of the scratch buffer before the swap. */
int b200geo_halo_block_in(const b200geo_grid *g, int member, int side, int kind, int width, int which,
                          void **ptr, uint64_t *bytes);
/* CUDA-IPC plumbing for direct NVLink P2P between one-process-per-GPU ranks. */
int b200geo_grid_ipc_export(const b200geo_grid *g, int which, void *handle64);
int b200geo_grid_ipc_open(b200geo_grid *g, int side, int which, const void *handle64);
/* Push this grid's owned boundary planes (all members, `width` planes) into the PEER ghost
 * planes of the neighbour opened on `side` with plain device-to-device copies over NVLink. */
int b200geo_halo_push(b200geo_grid *g, int side, int width, void *stream);
/* Tell the grid that `width` ghost planes on PEER side `side` of the current buffer are valid
 * (called after an exchange done by the caller, e.g. NCCL send/recv into b200geo_halo_block). */
int b200geo_halo_mark_valid(b200geo_grid *g, int side, int width);

/* ---- slab groups: one host thread, several GPUs of one box -------------------------------------
 * The slabs of ONE simulation space (b200geo_grid handles on different devices, in slab order along
 * the last axis; faces towards a neighbour are PEER ghost layers of equal width w) stepped together.
 * Replaces StripingSimulator::nanoStep (parallelization/stripingsimulator.h:269-286: rims first, ship
 * them, interior while they travel) and the PatchLink pair (communication/patchlink.h:127-151,218-244:
 * MPI_Isend/Irecv of packed regions) by direct NVLink copies of the contiguous rim planes into the
 * neighbours' ghost planes; one exchange per w sweeps (ghost zone width of
 * parallelization/nesting/vanillastepper.h:157-225), the w sweeps of a round fused into one launch
 * where the kernel family can (Jacobi). Called by B200StripingSimulator<CELL>
 * (include/libgeodecomp_b200/b200stripingsimulator.h). Does not own the grids. */
typedef struct b200geo_group b200geo_group;
int b200geo_group_create(b200geo_grid *const *grids, int n, int periodic, b200geo_group **out);
int b200geo_group_destroy(b200geo_group *grp);
/* the caller wrote cells (Initializer / Steerer): ghost planes are stale, exchange before stepping */
int b200geo_group_invalidate(b200geo_group *grp);
/* fill the w ghost planes of every PEER side of the current buffers now (whole cells) */
int b200geo_group_exchange(b200geo_group *grp);
/* n_steps x { UpdateFunctor over every slab; swap } with the exchanges they need; results are
 * bit-identical to b200geo_step on one grid holding the whole space */
int b200geo_group_step(b200geo_group *grp, int kernel, const void *params, uint32_t first_nano_step, uint32_t n_steps);
/* Same for cells whose update kernel lives OUTSIDE this library (the generic device path of
 * include/libgeodecomp_b200/b200generic.h: the user's own Cell::update compiled by nvcc): `update` is called with
 * the slab's device current and must enqueue, on `stream`, one sweep over the `dim` cells at `origin` (interior
 * coordinates of that slab), reading the grid's current buffer and writing its scratch buffer
 * (b200geo_grid_member_ptr which = 0 / 1); it returns 0 or a negative b200geo_status. One exchange of
 * ghost-width planes per sweep; rims first, interior overlapped with the transfer. */
typedef int (*b200geo_update_fn)(void *ctx, b200geo_grid *g, uint32_t nano_step, const int32_t origin[3],
                                 const int32_t dim[3], void *stream);
int b200geo_group_step_with(b200geo_group *grp, b200geo_update_fn update, void *ctx, uint32_t first_nano_step,
                            uint32_t n_steps);
int b200geo_group_sync(b200geo_group *grp);
/* out[0] = exchanges so far, out[1] = bytes shipped between devices */
int b200geo_group_stats(const b200geo_group *grp, uint64_t out[2]);

/* ---- BoxCell container grids: short-range n-body (B200GEO_KERNEL_NBODY) -------------------------
 * Replaces Grid<BoxCell<FixedArray<Particle, N> > > (storage/boxcell.h:21-178, storage/fixedarray.h:21-140)
 * and, on the step path, BoxCell::update with its re-binning (boxcell.h:112-174), the position checker
 * (misc/apitraits.h:1074-1088) and the particle iteration order of NeighborhoodIterator
 * (storage/neighborhooditerator.h:71-186) for the bound particle model (oracle/models/nbody.h).
 * Interchange format of a box of containers (what GridBase::set/get(Coord, BoxCell) carry, flattened):
 *   counts    int32 [dz][dy][dx]
 *   particles REAL  [dz][dy][dx][capacity][6]   pos x,y,z, vel x,y,z; slots >= count are ignored on
 *                                               load and zero on save
 * Container (x, y, z) covers [ (c + cell_origin) * cell_edge, + cell_edge ) per axis (doubles), which
 * is how the bound model's Initializer constructs BoxCell(origin, dimension). */
typedef struct {
    int32_t dim[3];           /* containers per axis */
    int32_t ghost_mode[3][2]; /* EDGE (Cube: the empty edge container) or, on the last axis, PEER */
    int32_t capacity;         /* N of FixedArray<Particle, N>, 1..64 */
    int32_t real_bytes;       /* 4: float particles, 8: double */
    int32_t cell_origin[3];   /* global index of container (0,0,0) (slab partitions) */
    double cell_edge;
} b200geo_boxgrid_desc;

typedef struct {
    double dt;
    double cutoff;            /* <= cell_edge */
    int32_t nano_steps;       /* APITraits::SelectNanoSteps of the cargo: re-bin when nanoStep % nano_steps == 0 */
} b200geo_nbody_params;

typedef struct b200geo_boxgrid b200geo_boxgrid;

int b200geo_boxgrid_create(const b200geo_boxgrid_desc *desc, int device, b200geo_boxgrid **out);
int b200geo_boxgrid_destroy(b200geo_boxgrid *g);
/* GridBase::set / get for a box of containers (ghost containers addressable with -1 / dim). */
int b200geo_boxgrid_load(b200geo_boxgrid *g, const int32_t origin[3], const int32_t dim[3], const int32_t *counts,
                         const void *particles, int location, int both, void *stream);
int b200geo_boxgrid_save(const b200geo_boxgrid *g, const int32_t origin[3], const int32_t dim[3], int32_t *counts,
                         void *particles, int location, void *stream);
/* n_steps x { re-bin (or copy), update every particle against the old grid; swap }. */
int b200geo_boxgrid_step(b200geo_boxgrid *g, const b200geo_nbody_params *params, uint32_t first_nano_step,
                         uint32_t n_steps, void *stream);
/* Synchronises and returns B200GEO_ERR_OUT_OF_RANGE ("capacity exceeded", storage/fixedarray.h:77-83)
 * if a container overflowed since the last check; the surplus particles were dropped. */
int b200geo_boxgrid_check(b200geo_boxgrid *g, void *stream);
/* One ghost (kind 1) or outermost owned (kind 0) plane of containers on `side` of the last axis, as a
 * contiguous block: array 0 = counts, 1 = particles; which = current (0) / scratch (1) buffer. */
int b200geo_boxgrid_halo_block(const b200geo_boxgrid *g, int array, int side, int kind, int which, void **ptr,
                               uint64_t *bytes);
int b200geo_boxgrid_halo_mark_valid(b200geo_boxgrid *g, int side, int width);

/* Slab group of container grids (slabs of containers along z, in order; faces towards a neighbour PEER): one
 * host thread, one slab per GPU. Per sweep every slab pulls its neighbours' boundary container planes (counts
 * and particles: one contiguous block each) straight into its ghost planes over NVLink, then re-bins / updates.
 * Replaces the MPI PatchLink pair for BoxCell grids; called by B200StripingSimulator<BoxCell<...> >. */
typedef struct b200geo_boxgroup b200geo_boxgroup;
int b200geo_boxgroup_create(b200geo_boxgrid *const *grids, int n, b200geo_boxgroup **out);
int b200geo_boxgroup_destroy(b200geo_boxgroup *grp);
int b200geo_boxgroup_step(b200geo_boxgroup *grp, const b200geo_nbody_params *params, uint32_t first_nano_step,
                          uint32_t n_steps);
int b200geo_boxgroup_sync(b200geo_boxgroup *grp);
/* out[0] = exchanges so far, out[1] = bytes shipped between devices */
int b200geo_boxgroup_stats(const b200geo_boxgroup *grp, uint64_t out[2]);

/* ---- ContainerCell grids: ID-keyed cargo (B200GEO_KERNEL_CONTAINER) ------------------------------
 * Replaces Grid<ContainerCell<CARGO, SIZE, int> > (storage/containercell.h:24-218) and, on the step path,
 * ContainerCell::update = copyOver + updateCargo (containercell.h:170-200) with the ID lookup of
 * NeighborhoodAdapter::operator[] (storage/neighborhoodadapter.h:45-65: the container itself first, then the other
 * containers of the 3^DIM Moore box in CoordBox order, x fastest; the first hit wins; each container searched by
 * upper_bound over its ascending ids, containercell.h:107-121) for the bound cargo model: the mesh element of
 * src/examples/voronoi/main.cpp:41-54 (oracle/models/container.h),
 *     temperature' = influx + (sum over neighborIDs, in list order, of hood[id].temperature) / neighborIDs.size().
 * The reference never adds or removes cargo during a run (containercell.h:44-46), so the engine resolves every
 * neighbour ID to a cargo index ONCE after a load and sweeps over resolved links from then on.
 * Interchange format of a box of containers (what GridBase::set/get(Coord, ContainerCell) carry, flattened; cells
 * in [dz][dy][dx] order, 2-D grids use dim[2] = 1):
 *   counts    int32  [cells]                            ContainerCell::size()
 *   ids       int32  [cells][capacity]                  ContainerCell::getIDs(): ascending
 *   values    double [cells][capacity]                  cargo.temperature
 *   influx    double [cells][capacity]                  cargo.influx
 *   nb_counts int32  [cells][capacity]                  cargo.neighborIDs.size()
 *   nb_ids    int32  [cells][capacity][max_neighbors]   cargo.neighborIDs
 * Slots >= count are ignored on load and zero on save. */
typedef struct {
    int32_t n_dims;           /* DIM of the cargo's topology, 2 or 3: the Moore box has 3^n_dims containers */
    int32_t dim[3];           /* containers per axis; dim[2] = 1 for n_dims = 2 */
    int32_t ghost_mode[3][2]; /* EDGE (Cube axis: lookups outside find the edge container) or WRAP (Torus axis), both sides alike */
    int32_t capacity;         /* SIZE of ContainerCell<CARGO, SIZE>, 1..4096 */
    int32_t max_neighbors;    /* capacity of the cargo's FixedArray<int, N> neighborIDs, 1..64 */
} b200geo_containergrid_desc;

typedef struct {
    int32_t *counts;
    int32_t *ids;
    double *values;
    double *influx;
    int32_t *nb_counts;
    int32_t *nb_ids;
} b200geo_container_box;

typedef struct b200geo_containergrid b200geo_containergrid;

int b200geo_containergrid_create(const b200geo_containergrid_desc *desc, int device, b200geo_containergrid **out);
int b200geo_containergrid_destroy(b200geo_containergrid *g);
/* GridBase::set for a box of containers: every array of `box` is read (location: where they live). */
int b200geo_containergrid_load(b200geo_containergrid *g, const int32_t origin[3], const int32_t dim[3],
                               const b200geo_container_box *box, int location, void *stream);
/* GridBase::get for a box of containers: arrays whose pointer is NULL are skipped (a Writer that only wants the
 * temperatures passes `values` alone). Returns when the data is in place. */
int b200geo_containergrid_save(const b200geo_containergrid *g, const int32_t origin[3], const int32_t dim[3],
                               const b200geo_container_box *box, int location, void *stream);
/* GridBase::setEdge / getEdge: one container (host arrays) found by lookups beyond a Cube boundary. */
int b200geo_containergrid_set_edge(b200geo_containergrid *g, const b200geo_container_box *cell);
int b200geo_containergrid_get_edge(const b200geo_containergrid *g, const b200geo_container_box *cell);
/* n_steps x { every cargo of every container: update against the old grid; swap }. The first call after a load
 * resolves the neighbour IDs: B200GEO_ERR_LOGIC ("id not found", neighborhoodadapter.h:63-64) if an ID a cargo
 * lists is in none of the 3^DIM containers around it, B200GEO_ERR_INVALID if a container's ids do not ascend,
 * B200GEO_ERR_OUT_OF_RANGE if a count exceeds its capacity; the grid is unchanged then. */
int b200geo_containergrid_step(b200geo_containergrid *g, uint32_t first_nano_step, uint32_t n_steps, void *stream);
/* out[0] = cargo items in the grid, out[1] = resolved neighbour links, out[2] = link resolutions so far,
 * out[3] = sweeps so far, out[4] = layout of the link table ("container.kernel": 0 or 1), out[5] = its bytes
 * (valid after the first step) */
int b200geo_containergrid_stats(const b200geo_containergrid *g, uint64_t out[6]);

/* ---- statistics: Simulator::gatherStatistics / Chronometer (misc/chronometer.h:142-150) ---- */
/* out[0] = device seconds spent in update kernels (TimeComputeInner), out[1] = seconds in ghost
 * refresh / halo copies (TimeComputeGhost + TimeCommunication), out[2] = number of sweeps. Uses
 * CUDA events recorded around b200geo_step when enabled. */
int b200geo_stats_enable(b200geo_grid *g, int on);
int b200geo_stats(b200geo_grid *g, double out[3]);

#ifdef __cplusplus
}
#endif

#endif
