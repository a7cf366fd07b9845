// Internal definitions shared by the translation units of libb200geo.so.
#ifndef B200GEO_CSRC_GRID_H
#define B200GEO_CSRC_GRID_H

#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#include "../../include/b200geo.h"

namespace b200geo {

// Device layout of one member array (HBM): [nz + 2gz][ny + 2gy][pitch] elements, x fastest.
// Every row starts on a 128-byte boundary and interior cell x = 0 sits exactly 128 bytes into
// the row, so 128-bit vector accesses to interior cells are aligned and a warp's 32 x 16 B
// requests coalesce into whole 128 B lines. Low-side x ghosts live at the end of the 128 B
// lead-in, high-side x ghosts directly behind the interior.
struct MemberLayout {
    int elem;            // bytes per element
    int lead;            // elements before interior x = 0 in a row (= 128 / elem)
    int64_t pitch;       // elements per row
    int64_t plane;       // elements per plane = pitch * (ny + 2gy)
    int64_t origin;      // element offset of interior (0,0,0)
    int64_t bytes;       // bytes of the whole member array (256 B aligned)
    int64_t offset;      // byte offset of the member inside a buffer
    int edge_offset;     // byte offset of this member inside an aggregated cell
};

struct Box {
    int x0, y0, z0, x1, y1, z1;  // half open, interior coordinates
};

}

struct b200geo_grid {
    b200geo_grid_desc desc;
    int device;
    int n, g[3], d[3];
    int slab_axis;       // axis slabs are cut along: the last used axis (2, or 1 for 2-D grids)
    int64_t uniform_stride; // elements per member array in the uniform element layout (0 = default layout)
    b200geo::MemberLayout m[B200GEO_MAX_MEMBERS];
    int64_t buffer_bytes;
    char *buf[2];        // buf[cur] = current, buf[cur ^ 1] = scratch
    int cur;
    int cell_bytes;
    unsigned char edge[8 * B200GEO_MAX_MEMBERS];
    int peer_valid[2];   // valid ghost planes per z side (PEER sides)
    char *peer_buf[2][2];   // [side][which]: IPC-mapped buffers of the z neighbours
    bool stats_on;
    cudaEvent_t ev[4];
    double t_update, t_ghost;
    uint64_t sweeps;
    void *bits;          // bit-packed copy of a Game of Life grid (two buffers), allocated on first use
    size_t bits_bytes;
    void *scratch;       // small device scratch (streak lists etc.)
    size_t scratch_bytes;
    char *io_staging;    // device staging of host-side region I/O (grow-only: no cudaMalloc per set/get call)
    size_t io_staging_bytes;

    char *member_ptr(int member, int which) const { return buf[cur ^ which] + m[member].offset; }
    // elements per slice (plane, or row for 2-D grids) along the slab axis
    int64_t slice_elems(int member) const { return slab_axis == 2 ? m[member].plane : m[member].pitch; }
};

namespace b200geo {

struct Tuning {
    int jacobi_zchunk;    // planes per CTA along z (0 = automatic)
    int jacobi_prefetch;  // L2 prefetch distance in planes (0 = off, < 0 = per-kernel default)
    int gol_rows;         // rows per CTA (0 = automatic)
    int lbm_block;        // threads per CTA
    int jacobi_tb;        // temporal blocking depth of the Jacobi kernels (sweeps per launch; 0 = automatic, 1 = off)
    int jacobi_tb_rows;   // tile shape of the temporal-blocked kernel: 32 (2 rows/thread), 64 or 33 (4 rows/thread)
    int jacobi_tb_zchunk; // planes per CTA along z of the temporal-blocked kernel (0 = automatic)
    int gol_bits;         // fewest sweeps per b200geo_step call for which Game of Life runs bit-packed (0 = never)
    int gol_bits_rows;    // rows per CTA of the bit-packed kernel (0 = automatic)
    int nbody_kernel;     // 1 = re-bin kernel + one-pass force kernel, 3 = fused kernel with candidate lists (default), 0 = the fused kernel with per-container masks (measured 10 % slower, profiles/r3d)
    int jacobi_pdl;       // programmatic dependent launch of the one-sweep Jacobi kernel: 0 / 1, < 0 = small grids only
    int lbm_variant;      // rows per thread of the LBM kernel: 1 (default) or 2
    int jacobi_tb_promo;  // L2 promotion of the temporal-blocked kernel's TMA loads: 0 none (default: 5 % faster, 9 % fewer DRAM reads than 256 B, profiles/r3c), 1 64 B, 2 128 B, 3 256 B
    int nbody_run;        // containers per CTA of the fused n-body kernel: 8, 12 or 16 (0 = automatic: 16 for float, 8 for double)
    int nbody_threads;    // threads per CTA of the fused n-body kernel (0 = automatic: 16 per container of a run)
    int jacobi_resident;  // SM-resident multi-sweep kernel for small grids: 1 = whenever the grid qualifies, 0 = never (default: not faster, profiles/r3i_r3j_r3k)
    int jacobi_tb_raster; // CTA order of the temporal-blocked kernel: 0 = x fastest (default), n > 0 = y-panels of n tile rows, y fastest
    int lbm_tb;           // sweeps per launch of the LBM kernels: 2 = two fused sweeps wherever the ghost zones allow (default), 1 = one
    int lbm_tb_rows;      // rows of the intermediate level per CTA of the fused LBM kernel: 14 (default), 16 or 8
    int lbm_tb_zchunk;    // planes per CTA along z of the fused LBM kernel (0 = automatic)
    int lbm_tb_promo;     // L2 promotion of the fused LBM kernel's TMA loads: 0 none, 1 64 B, 2 128 B, 3 256 B
    int lbm_tb_warps;     // fused LBM kernel: 1 = sweep 1 and sweep 2 on different warps linked by mbarriers, 0 = all warps do both, two CTA-wide barriers per plane
    int lbm_tb_hints;     // L2 hints of the fused LBM kernel: bit 0 = streaming stores, bit 1 = evict-last window loads
    int container_kernel; // ContainerCell sweeps: 0 = links into the global value array (windows of 1024 sorted by neighbour count), 1 = tiles of containers with their neighbourhood's values staged in shared memory and 16-bit links (where the capacity allows)
};
extern Tuning g_tuning;

int fail(int status, const std::string& msg);
int check_cuda(cudaError_t e, const char *what);
void count_launch(uint64_t n = 1);

// kernel families (one translation unit each)
int sweep_jacobi(b200geo_grid *g, int kind, const Box& box, cudaStream_t s);
int sweep_jacobi_tb(b200geo_grid *g, int kind, int depth, const Box& box, cudaStream_t s);
int jacobi_resident_planes(const b200geo_grid *g, int kind);
int sweep_jacobi_resident(b200geo_grid *g, int kind, int planes, int sweeps, cudaStream_t s);
int sweep_gol(b200geo_grid *g, const Box& box, cudaStream_t s);
bool gol_bits_applicable(const b200geo_grid *g);
int sweep_gol_bits(b200geo_grid *g, uint32_t sweeps, cudaStream_t s);
int sweep_lbm(b200geo_grid *g, const Box& box, bool store_macroscopic, cudaStream_t s);
int sweep_lbm_tb2(b200geo_grid *g, const Box& box, bool store_macroscopic, cudaStream_t s);

// ghost maintenance and (de)serialisation kernels (region.cu)
int fill_edge(b200geo_grid *g, int which, cudaStream_t s);
int refresh_wrap(b200geo_grid *g, cudaStream_t s);
int copy_region(b200geo_grid *g, const int32_t *streaks, int n_streaks, char *dev_buf, int64_t count,
                bool save, int which, cudaStream_t s);

#define B200GEO_CUDA(call)                                                    \
    do {                                                                      \
        int b200geo_rc_ = ::b200geo::check_cuda((call), #call);               \
        if (b200geo_rc_ != 0) return b200geo_rc_;                             \
    } while (0)

}

#endif
