// SM-resident Jacobi 6/7-point kernel for small grids (BASELINE.json configs[0]: 128^3, 100 steps).
//
// A grid of a few million cells fits the shared memory and the register files of the whole chip
// (128^3 f64 = 16.8 MB; 148 SMs x 227 KB = 33.6 MB). The one-sweep kernel (jacobi.cu) streams such a grid
// out of L2 and back once per sweep and pays a launch per sweep: 7-10 us, L2-bandwidth and launch bound.
// Here ONE cooperative launch does all the sweeps of a b200geo_step call: every CTA owns a brick of
// nx x 16 x BZ cells for the whole call, keeps it in registers (a thread owns two x-adjacent cells of two
// rows of every plane of the brick) with a copy in shared memory for its neighbours inside the CTA, and per
// sweep only the brick's FACES travel: they are written to their true places in the scratch grid buffer, a
// per-brick flag (release / acquire, no grid-wide barrier: a brick waits for its four face neighbours only) makes
// them visible, the neighbouring bricks read them into the one-cell shell around their own brick. Per sweep and
// CTA 48 KB out and 53 KB in instead of the 2 x 128 KB of the whole brick, and no launch. The grid buffers alternate like the sweeps of the other kernels do, the
// last sweep stores the whole brick; what the scratch buffer holds afterwards is unspecified, as always.
//
// Same arithmetic per cell as jacobi.cu (and therefore as the reference's Cell::update): bit-identical.
// Cube topologies (the constant edge cell is read out of the ghost ring, which both buffers hold), nx <= 128
// and even, ny a multiple of 16, nz a multiple of BZ, at most one brick per SM; everything else takes the
// streaming kernels.
#include "grid.h"

namespace b200geo {

namespace {

constexpr int RX = 64;          // x pairs per row
constexpr int RY = 8;           // thread rows; a thread owns rows 2 ty and 2 ty + 1
constexpr int BY = 2 * RY;      // rows of a brick
constexpr int NT = RX * RY;     // 512 threads
constexpr int SP = 132;         // shared-memory row pitch in doubles: x = -1 at 1, x = 0 at 2 (16-byte aligned)

template<int BZ>
__device__ __forceinline__ int sidx(int x, int y, int z)
{
    return ((z + 1) * (BY + 2) + (y + 1)) * SP + x + 2;
}

__device__ __forceinline__ void flag_release(int *flag, int value)
{
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
}

__device__ __forceinline__ int flag_acquire(const int *flag)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    return v;
}

__device__ __forceinline__ double2 lds2(const double *p)
{
    return *reinterpret_cast<const double2 *>(p);
}

template<int KIND, int BZ>
__global__ void __launch_bounds__(NT, 1)
jacobi_resident_kernel(double *buf_a, double *buf_b, int64_t pitch, int64_t plane, int nx, int nby, int nbz, int sweeps, int *flags)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *s = reinterpret_cast<double *>(smem_raw);

    const int tx = threadIdx.x % RX, ty = threadIdx.x / RX;
    const int x = 2 * tx, y0 = 2 * ty;
    const int Y0 = (blockIdx.x % nby) * BY, Z0 = (blockIdx.x / nby) * BZ;
    const bool live = x < nx;
    // the face neighbours of this brick (y - 1, y + 1, z - 1, z + 1), -1 = the domain boundary
    int neighbour = -1;
    if (threadIdx.x < 4) {
        const int by = blockIdx.x % nby, bz = blockIdx.x / nby;
        const int ny_ = by + (threadIdx.x == 0 ? -1 : threadIdx.x == 1 ? 1 : 0), nz_ = bz + (threadIdx.x == 2 ? -1 : threadIdx.x == 3 ? 1 : 0);
        if (ny_ >= 0 && ny_ < nby && nz_ >= 0 && nz_ < nbz) neighbour = nz_ * nby + ny_;
    }
    const int64_t base = (int64_t)Z0 * plane + (int64_t)Y0 * pitch;

    // the brick and its shell (ghost ring / neighbouring bricks) at time 0, x ghosts included
    for (int row = threadIdx.x / 32; row < (BZ + 2) * (BY + 2); row += NT / 32) {
        const int z = row / (BY + 2) - 1, y = row % (BY + 2) - 1;
        const double *g = buf_a + base + (int64_t)z * plane + (int64_t)y * pitch;
        for (int i = (int)(threadIdx.x % 32) - 1; i <= nx; i += 32) s[sidx<BZ>(i, y, z)] = __ldcg(g + i);
    }
    __syncthreads();
    double2 c[BZ][2];
#pragma unroll
    for (int k = 0; k < BZ; ++k)
#pragma unroll
        for (int r = 0; r < 2; ++r) c[k][r] = live ? lds2(s + sidx<BZ>(x, y0 + r, k)) : make_double2(0.0, 0.0);

    // the shell pairs: (offset in the grid buffer relative to the brick, offset in shared memory)
    int2 *shell = reinterpret_cast<int2 *>(s + (BZ + 2) * (BY + 2) * SP);
    const int shell_pairs = (2 * (BZ + 2) + 2 * BY) * (nx / 2);
    for (int i = threadIdx.x; i < shell_pairs; i += NT) {
        const int pairs = nx / 2, row = i / pairs, px = 2 * (i % pairs);
        int y, z;
        if (row < 2 * (BZ + 2)) {
            z = row / 2 - 1;
            y = (row & 1) ? BY : -1;
        } else {
            const int q = row - 2 * (BZ + 2);
            z = (q & 1) ? BZ : -1;
            y = q / 2;
        }
        shell[i] = make_int2((int)(z * plane + y * pitch + px), sidx<BZ>(px, y, z));
    }
    __syncthreads();

    for (int t = 0; t < sweeps; ++t) {
        // ---- one sweep over the brick, in place: z neighbours are the thread's own registers
        if (live) {
            double2 below[2];
            below[0] = lds2(s + sidx<BZ>(x, y0, -1));
            below[1] = lds2(s + sidx<BZ>(x, y0 + 1, -1));
#pragma unroll
            for (int k = 0; k < BZ; ++k) {
                double2 old[2] = {c[k][0], c[k][1]};
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const double2 zm = below[r];
                    const double2 zp = k == BZ - 1 ? lds2(s + sidx<BZ>(x, y0 + r, BZ)) : c[k == BZ - 1 ? k : k + 1][r];
                    const double2 ym = r == 0 ? lds2(s + sidx<BZ>(x, y0 - 1, k)) : old[0];
                    const double2 yp = r == 1 ? lds2(s + sidx<BZ>(x, y0 + 2, k)) : old[1];
                    const double w = s[sidx<BZ>(x - 1, y0 + r, k)], e = s[sidx<BZ>(x + 2, y0 + r, k)];
                    const double2 m = old[r];
                    double2 v;
                    if (KIND == 6) {
                        v.x = (zm.x + ym.x + w + m.y + yp.x + zp.x) * (1.0 / 6.0);
                        v.y = (zm.y + ym.y + m.x + e + yp.y + zp.y) * (1.0 / 6.0);
                    } else {
                        v.x = (zm.x + ym.x + w + m.x + m.y + yp.x + zp.x) * (1.0 / 7.0);
                        v.y = (zm.y + ym.y + m.x + m.y + e + yp.y + zp.y) * (1.0 / 7.0);
                    }
                    c[k][r] = v;
                }
                below[0] = old[0];
                below[1] = old[1];
            }
        }
        const bool last = t + 1 == sweeps;
        // ---- the faces (the last sweep: the whole brick) to their places in the other grid buffer
        if (live) {
            double *g = buf_b + base + x;
#pragma unroll
            for (int k = 0; k < BZ; ++k)
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const bool face = k == 0 || k == BZ - 1 || (ty == 0 && r == 0) || (ty == RY - 1 && r == 1);
                    if (last || face) *reinterpret_cast<double2 *>(g + (int64_t)k * plane + (int64_t)(y0 + r) * pitch) = c[k][r];
                }
        }
        if (last) break;
        __syncthreads();      // every thread's face cells are ordered before the barrier; everybody has read the old brick
        // release at gpu scope is cumulative: it publishes the face cells of the whole CTA (they happen before the barrier)
        if (threadIdx.x == 0) flag_release(flags + blockIdx.x, t + 1);
        if (live) {
#pragma unroll
            for (int k = 0; k < BZ; ++k)
#pragma unroll
                for (int r = 0; r < 2; ++r) *reinterpret_cast<double2 *>(s + sidx<BZ>(x, y0 + r, k)) = c[k][r];
        }
        // wait for the faces of the four neighbours (they cannot run ahead by more than one sweep: they wait for ours)
        if (neighbour >= 0) {
            while (flag_acquire(flags + neighbour) < t + 1) {}
        }
        __syncthreads();
        // ---- the shell: rows y = -1 and y = BY of every plane, planes z = -1 and z = BZ (L2, not L1: the same
        // addresses held another time step two sweeps ago); where each pair lies was worked out once, before the sweeps
        for (int i = threadIdx.x; i < shell_pairs; i += NT) {
            const int2 at = shell[i];
            const double2 v = __ldcg(reinterpret_cast<const double2 *>(buf_b + base + at.x));
            *reinterpret_cast<double2 *>(s + at.y) = v;
        }
        __syncthreads();
        double *tmp = buf_a;
        buf_a = buf_b;
        buf_b = tmp;
    }
}

template<int KIND, int BZ>
int launch_resident(b200geo_grid *g, int sweeps, cudaStream_t s)
{
    const MemberLayout& L = g->m[0];
    double *a = (double *)g->member_ptr(0, 0) + L.origin;
    double *b = (double *)g->member_ptr(0, 1) + L.origin;
    int64_t pitch = L.pitch, plane = L.plane;
    int nx = g->d[0], nby = g->d[1] / BY, nbz = g->d[2] / BZ;
    int bricks = nby * nbz;
    // one flag per brick: the time step whose faces it has published
    if ((size_t)bricks * sizeof(int) > g->scratch_bytes) {
        if (g->scratch) cudaFree(g->scratch);
        g->scratch = 0;
        g->scratch_bytes = 0;
        B200GEO_CUDA(cudaMalloc(&g->scratch, 4096));
        g->scratch_bytes = 4096;
    }
    int *flags = (int *)g->scratch;
    B200GEO_CUDA(cudaMemsetAsync(flags, 0, (size_t)bricks * sizeof(int), s));
    size_t smem = (size_t)(BZ + 2) * (BY + 2) * SP * sizeof(double) + (size_t)(2 * (BZ + 2) + 2 * BY) * RX * sizeof(int2);
    auto kernel = jacobi_resident_kernel<KIND, BZ>;
    static bool attr_set[64] = {false};
    if (g->device < 0 || g->device >= 64 || !attr_set[g->device]) {
        B200GEO_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (g->device >= 0 && g->device < 64) attr_set[g->device] = true;
    }
    void *args[] = {&a, &b, &pitch, &plane, &nx, &nby, &nbz, &sweeps, &flags};
    B200GEO_CUDA(cudaLaunchCooperativeKernel((void *)kernel, dim3(bricks), dim3(NT), args, smem, s));
    count_launch();
    return check_cuda(cudaGetLastError(), "resident jacobi sweeps");
}

int sm_count(int device)
{
    static int cached[64] = {0};
    if (device >= 0 && device < 64 && cached[device]) return cached[device];
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    if (device >= 0 && device < 64) cached[device] = n;
    return n;
}

}

// planes per brick (8 or 4) if the SM-resident kernel can take this grid, else 0
int jacobi_resident_planes(const b200geo_grid *g, int kind)
{
    if (kind != 6 && kind != 7) return 0;
    if (g->n != 1 || g->m[0].elem != 8) return 0;
    for (int i = 0; i < 3; ++i) {
        if (g->g[i] < 1) return 0;
        for (int side = 0; side < 2; ++side)
            if (g->desc.ghost_mode[i][side] != B200GEO_GHOST_EDGE) return 0;
    }
    if (g->d[0] > 2 * RX || g->d[0] % 2 || g->d[1] % BY) return 0;
    const int sms = sm_count(g->device);
    const int want[2] = {8, 4};
    for (int i = 0; i < 2; ++i) {
        const int bz = want[i];
        if (g->d[2] % bz) continue;
        const int bricks = (g->d[1] / BY) * (g->d[2] / bz);
        // one brick per SM at most (cooperative launch), and not fewer than half the machine
        if (bricks <= sms && 2 * bricks >= sms) return bz;
    }
    return 0;
}

int sweep_jacobi_resident(b200geo_grid *g, int kind, int planes, int sweeps, cudaStream_t s)
{
    if (planes == 8) return kind == 6 ? launch_resident<6, 8>(g, sweeps, s) : launch_resident<7, 8>(g, sweeps, s);
    return kind == 6 ? launch_resident<6, 4>(g, sweeps, s) : launch_resident<7, 4>(g, sweeps, s);
}

}
