/* Generic device path for UNBOUND user cells: the user's own Cell::update() — or
 * the AoS-signature Cell::updateLineX() — compiled by nvcc into a sm_100a kernel of the user's
 * translation unit, one sweep per launch, on the device grid of libb200geo.so. Any model whose
 * update is `__host__ __device__` (what the reference's CUDASimulator requires as well,
 * parallelization/cudasimulator.h:160-236) runs on B200Simulator<CELL> without a hand-written
 * kernel and without a B200GEO_BIND_CELL line; bound cells keep their hand kernels.
 *
 * Included by b200simulator.h when the translation unit is compiled by nvcc. Differences to the
 * reference's CUDA path, on purpose:
 *  - the cell is stored word-sliced: a CELL of N machine words is N SoA members of the device
 *    grid, so a warp's access to word w of 32 x-adjacent cells is one coalesced request (the
 *    reference keeps AoS cells on the device, storage/cudagrid.h). The neighbourhood hands out
 *    cells BY VALUE, assembled from their words; loads of words the model never reads are dead
 *    code the compiler removes.
 *  - neighbours may be addressed with FixedCoord<X,Y,Z> AND with run-time Coord<DIM> (the
 *    reference's HoodType accepts FixedCoord only, cudasimulator.h:52-64, so e.g. the Game of Life
 *    example cannot run there): edge cells and periodic images are real cells in the grid's ghost
 *    ring (EDGE / WRAP layers), no boundary arithmetic per access.
 *  - `*this` inside update() is the cell of the NEW grid, exactly as in VanillaUpdateFunctor
 *    (storage/vanillaupdatefunctor.h:28-32): a model that does not assign every member sees the
 *    value from two sweeps ago, like on the CPU.
 * Struct-of-Arrays cells (APITraits::HasSoA + HasUpdateLineX) take the sibling path of b200genericsoa.h:
 * their SoA-signature updateLineX(hoodOld, indexEnd, hoodNew, nanoStep) runs with the accessors
 * LIBFLATARRAY_REGISTER_SOA generated for them. Not covered: static data (APITraits::HasStaticData),
 * word-sliced cells larger than 32 words.
 */
#ifndef LIBGEODECOMP_B200_B200GENERIC_H
#define LIBGEODECOMP_B200_B200GENERIC_H

#ifndef __CUDACC__
#error "b200generic.h needs nvcc: the user's update() is compiled into a device kernel"
#endif

#include <cuda_runtime.h>

#include <libgeodecomp/geometry/coord.h>
#include <libgeodecomp/geometry/fixedcoord.h>
#include <libgeodecomp/misc/apitraits.h>

#include <cstring>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

/* kernel id of cells without a hand-written kernel: the step loop lives in the user's TU */
#define B200GEO_KERNEL_GENERIC 0

namespace LibGeoDecomp {

namespace B200Generic {

template<int W> struct WordType;
template<> struct WordType<8> { typedef unsigned long long Type; };
template<> struct WordType<4> { typedef unsigned int Type; };
template<> struct WordType<2> { typedef unsigned short Type; };
template<> struct WordType<1> { typedef unsigned char Type; };

/* how a CELL is sliced into words: the widest word its size and alignment allow */
template<typename CELL>
struct Words {
    static const int W =
        (sizeof(CELL) % 8 == 0 && alignof(CELL) % 8 == 0) ? 8 :
        (sizeof(CELL) % 4 == 0 && alignof(CELL) % 4 == 0) ? 4 :
        (sizeof(CELL) % 2 == 0 && alignof(CELL) % 2 == 0) ? 2 : 1;
    static const int N = sizeof(CELL) / W;
    typedef typename WordType<W>::Type Word;
};

/* one of the two buffers: word w of cell i lives at base + w * memberStride + i * sizeof(Word) */
struct View {
    char *base;
    long long memberStride;  /* bytes between two word arrays */
    long long pitch, plane;  /* cells per row / per plane of the padded arrays */
};

/* raw bytes of one cell: user cells need not have a __device__ default constructor (the reference's
 * TestCell has none); a cell is materialised by copying its words, then handed out by copy */
template<typename CELL>
struct Storage {
    alignas(CELL) alignas(8) unsigned char raw[sizeof(CELL)];

    __device__ __forceinline__ CELL *cell()
    {
        return reinterpret_cast<CELL*>(raw);
    }
};

template<typename CELL>
__device__ __forceinline__ void gather(CELL *cell, const View& v, long long index)
{
    typedef typename Words<CELL>::Word Word;
#pragma unroll
    for (int w = 0; w < Words<CELL>::N; ++w) {
        Word word = *reinterpret_cast<const Word*>(v.base + w * v.memberStride + index * (long long)sizeof(Word));
        memcpy(reinterpret_cast<char*>(cell) + w * sizeof(Word), &word, sizeof(Word));
    }
}

template<typename CELL>
__device__ __forceinline__ void scatter(const CELL& cell, const View& v, long long index)
{
    typedef typename Words<CELL>::Word Word;
#pragma unroll
    for (int w = 0; w < Words<CELL>::N; ++w) {
        Word word;
        memcpy(&word, reinterpret_cast<const char*>(&cell) + w * sizeof(Word), sizeof(Word));
        *reinterpret_cast<Word*>(v.base + w * v.memberStride + index * (long long)sizeof(Word)) = word;
    }
}

/* The neighbourhood object a model's update() receives on the device. Plays the role of
 * CoordMap (storage/coordmap.h:34-43), FixedNeighborhood (storage/fixedneighborhood.h:66-83) and
 * LinePointerNeighborhood: relative addressing around the cell being updated, offset by *x for
 * updateLineX-style models that advance an index. */
template<typename CELL, int DIM>
class Hood
{
public:
    __device__ Hood(const View& view, long long center, const long *x) :
        view(view),
        center(center),
        x(x)
    {}

    template<int X, int Y, int Z>
    __device__ __forceinline__ CELL operator[](FixedCoord<X, Y, Z>) const
    {
        Storage<CELL> s;
        gather(s.cell(), view, center + *x + X + Y * view.pitch + Z * view.plane);
        return *s.cell();
    }

    __device__ __forceinline__ CELL operator[](const Coord<1>& c) const
    {
        Storage<CELL> s;
        gather(s.cell(), view, center + *x + c.x());
        return *s.cell();
    }

    __device__ __forceinline__ CELL operator[](const Coord<2>& c) const
    {
        Storage<CELL> s;
        gather(s.cell(), view, center + *x + c.x() + c.y() * view.pitch);
        return *s.cell();
    }

    __device__ __forceinline__ CELL operator[](const Coord<3>& c) const
    {
        Storage<CELL> s;
        gather(s.cell(), view, center + *x + c.x() + c.y() * view.pitch + c.z() * view.plane);
        return *s.cell();
    }

private:
    View view;
    long long center;
    const long *x;
};

/* does `cell.update(hood, nanoStep)` compile? (TestCell advertises updateLineX too, but only its
 * update() is __host__ __device__, misc/testcell.h:195-233 — update() wins when both exist) */
template<typename CELL, typename HOOD>
struct HasUpdate {
    template<typename C>
    static char test(decltype(std::declval<C&>().update(std::declval<const HOOD&>(), 0u), 0) *);
    template<typename C>
    static long test(...);
    static const bool VALUE = sizeof(test<CELL>(0)) == sizeof(char);
};

template<typename CELL, typename HOOD, bool HAS_UPDATE>
struct Invoke;

template<typename CELL, typename HOOD>
struct Invoke<CELL, HOOD, true> {
    __device__ static void run(CELL *cell, long *, const HOOD& hood, unsigned nanoStep)
    {
        cell->update(hood, nanoStep);
    }
};

/* AoS line signature, storage/linepointerupdatefunctor.h:170-179: a line of length one */
template<typename CELL, typename HOOD>
struct Invoke<CELL, HOOD, false> {
    __device__ static void run(CELL *cell, long *x, const HOOD& hood, unsigned nanoStep)
    {
        CELL::updateLineX(cell, x, 1, hood, nanoStep);
    }
};

/* One sweep over a box: thread = cell, x fastest (coalesced per word array), y and z from the block index.
 * `origin` = element index of the box's first cell. */
template<typename CELL, int DIM>
__global__ void __launch_bounds__(256)
updateKernel(View oldView, View newView, long long origin, int nx, int ny, unsigned nanoStep)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int z = blockIdx.z;
    if (x >= nx || y >= ny) {
        return;
    }
    const long long index = origin + x + y * oldView.pitch + z * oldView.plane;
    typedef Hood<CELL, DIM> HoodType;
    Storage<CELL> cell;
    gather(cell.cell(), newView, index);
    long lineX = 0;
    HoodType hood(oldView, index, &lineX);
    Invoke<CELL, HoodType, HasUpdate<CELL, HoodType>::VALUE>::run(cell.cell(), &lineX, hood, nanoStep);
    scatter(*cell.cell(), newView, index);
}

inline void check(int rc)
{
    if (rc < 0) {
        std::string msg = b200geo_last_error();
        throw std::runtime_error(msg.find("CUDA error") == 0 ? msg : "CUDA error: " + msg);
    }
}

template<typename CELL>
inline View view(b200geo_grid *g, int which)
{
    View v;
    void *p0 = 0, *p1 = 0;
    int64_t pitch = 0, plane = 0;
    check(b200geo_grid_member_ptr(g, 0, which, &p0));
    check(b200geo_grid_layout(g, 0, &pitch, &plane, 0));
    v.base = static_cast<char*>(p0);
    v.memberStride = 0;
    if (Words<CELL>::N > 1) {
        check(b200geo_grid_member_ptr(g, 1, which, &p1));
        v.memberStride = static_cast<char*>(p1) - static_cast<char*>(p0);
    }
    v.pitch = pitch;
    v.plane = plane;
    return v;
}

/* the word-sliced binding: any cell with a __host__ __device__ update() or AoS-signature updateLineX() */
template<typename CELL>
struct WordSlicedBinding {
    typedef typename APITraits::SelectTopology<CELL>::Value Topology;
    static const int DIM = Topology::DIM;
    static const unsigned NANO_STEPS = APITraits::SelectNanoSteps<CELL>::VALUE;

    static_assert(B200Generic::Words<CELL>::N <= B200GEO_MAX_MEMBERS,
                  "generic B200 path: cell larger than 32 machine words; bind a hand-written kernel or shrink the cell");

    static int kernel()
    {
        return B200GEO_KERNEL_GENERIC;
    }

    /* the word table: word w = bytes [w * W, (w + 1) * W) of the cell */
    static std::vector<B200Member> members()
    {
        std::vector<B200Member> ret(B200Generic::Words<CELL>::N);
        for (int w = 0; w < B200Generic::Words<CELL>::N; ++w) {
            ret[w].offsetInCell = (std::size_t)w * B200Generic::Words<CELL>::W;
            ret[w].bytes = B200Generic::Words<CELL>::W;
        }
        return ret;
    }

    /* enqueue one sweep over the box (origin, dim) of grid g on `stream`: current buffer -> scratch buffer */
    static void launchBox(b200geo_grid *g, unsigned nanoStep, const int32_t origin[3], const int32_t dim[3], cudaStream_t stream)
    {
        if (dim[0] <= 0 || dim[1] <= 0 || dim[2] <= 0) {
            return;
        }
        int64_t first = 0;
        B200Generic::check(b200geo_grid_layout(g, 0, 0, 0, &first));
        B200Generic::View oldView = B200Generic::view<CELL>(g, 0);
        B200Generic::View newView = B200Generic::view<CELL>(g, 1);
        first += origin[0] + origin[1] * oldView.pitch + origin[2] * oldView.plane;
        dim3 block(128, DIM > 1 ? 2 : 1, 1);
        dim3 grid((dim[0] + block.x - 1) / block.x, (dim[1] + block.y - 1) / block.y, dim[2]);
        if (grid.y > 65535u || grid.z > 65535u) {
            throw std::out_of_range("grid dimension too large");
        }
        B200Generic::updateKernel<CELL, DIM><<<grid, block, 0, stream>>>(
            oldView, newView, (long long)first, dim[0], dim[1], nanoStep % NANO_STEPS);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) {
            throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(e) + " in generic update kernel");
        }
    }

    /* one sweep over a box: current buffer -> scratch buffer, no swap (B200Grid::updateRegion, B200Stepper) */
    static void updateBox(b200geo_grid *g, unsigned nanoStep, const int32_t origin[3], const int32_t dim[3])
    {
        launchBox(g, nanoStep, origin, dim, 0);
    }

    /* sweeps x { refresh periodic images; UpdateFunctor over the whole grid; swap }
     * = SerialSimulator::nanoStep (parallelization/serialsimulator.h:132-139) */
    static void step(b200geo_grid *g, const int32_t dim[3], unsigned firstNanoStep, unsigned sweeps)
    {
        const int32_t origin[3] = {0, 0, 0};
        for (unsigned t = 0; t < sweeps; ++t) {
            B200Generic::check(b200geo_refresh_ghosts(g, 0));
            launchBox(g, firstNanoStep + t, origin, dim, 0);
            B200Generic::check(b200geo_swap(g));
        }
    }

    /* the same on a slab group (B200StripingSimulator): the group drives the schedule — rims, halo copies,
     * interiors — and calls back for every box it wants updated */
    static int updateCallback(void *, b200geo_grid *g, uint32_t nanoStep, const int32_t origin[3], const int32_t dim[3], void *stream)
    {
        try {
            int device = 0;
            B200Generic::check(b200geo_grid_device(g, &device));
            if (cudaSetDevice(device) != cudaSuccess) {
                return B200GEO_ERR_CUDA;
            }
            launchBox(g, nanoStep, origin, dim, static_cast<cudaStream_t>(stream));
        } catch (const std::exception&) {
            return B200GEO_ERR_CUDA;
        }
        return B200GEO_OK;
    }

    static void groupStep(b200geo_group *group, unsigned firstNanoStep, unsigned sweeps)
    {
        B200Generic::check(b200geo_group_step_with(group, &updateCallback, 0, firstNanoStep, sweeps));
    }
};

/* which generic binding a cell gets: Struct-of-Arrays cells whose only update is a SoA-signature updateLineX() are
 * handed the accessors LibFlatArray generated for them (b200genericsoa.h) — the route UpdateFunctor takes for them on
 * the CPU (storage/updatefunctor.h:403-428 -> FixedNeighborhoodUpdateFunctor). Every other cell is word-sliced; that
 * includes SoA cells that ALSO have a per-cell update(), which wins as it does for AoS cells (the reference's
 * TestCellSoA is such a cell, and only its update() is __host__ __device__, misc/testcell.h:213-281). */
template<typename CELL, typename SOA = typename APITraits::SelectSoA<CELL>::Value,
         typename LINE = typename APITraits::SelectUpdateLineX<CELL>::Value,
         bool HAS_UPDATE = HasUpdate<CELL, Hood<CELL, APITraits::SelectTopology<CELL>::Value::DIM> >::VALUE>
struct SelectBinding {
    typedef WordSlicedBinding<CELL> Type;
};

template<typename CELL>
struct SelectBinding<CELL, APITraits::TrueType, APITraits::TrueType, false> {
    typedef SoA::Binding<CELL, SoA::DeviceSweep> Type;
};

}

/* Primary template = the generic path. B200GEO_BIND_CELL specialises it for bound cells. */
template<typename CELL>
struct B200KernelBinding : public B200Generic::SelectBinding<CELL>::Type
{};

}

#endif
