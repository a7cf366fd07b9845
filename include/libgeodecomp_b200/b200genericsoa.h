/* Generic device path for UNBOUND Struct-of-Arrays cells (second half): a model's
 * SoA-signature
 *
 *     template<typename HOOD_OLD, typename HOOD_NEW>
 *     static void updateLineX(HOOD_OLD& hoodOld, int indexEnd, HOOD_NEW& hoodNew, int nanoStep)
 *
 * (what FixedNeighborhoodUpdateFunctor calls, storage/fixedneighborhoodupdatefunctor.h:126-256; e.g.
 * src/testbed/performancetests/main.cpp:1303-1395) compiled by nvcc into a sm_100a kernel of the user's translation
 * unit, on the device grid of libb200geo.so. The member functions a model uses on its hoods — `hoodNew.temp()`,
 * `hoodOld[FixedCoord<0, 0, -1>()].temp()` — exist only on the accessor classes LIBFLATARRAY_REGISTER_SOA generates
 * for the cell (lib/libflatarray/include/libflatarray/macros.hpp:129-390), so those very classes are used here:
 *
 *   - the grid is created in the engine's UNIFORM ELEMENT LAYOUT (b200geo_grid_create_uniform): one element index
 *     addresses every member of a cell and member m starts DIM_PROD x offset<CELL, m> bytes into the buffer, which
 *     is LibFlatArray's addressing contract (macros.hpp:327-349). Rows keep their 128-byte alignment and the ghost
 *     ring (EDGE / WRAP / PEER layers) keeps working unchanged, halos included.
 *   - LibFlatArray fixes the grid extents at compile time (DIM_X x DIM_Y x DIM_Z from a list of cubes, with a
 *     run-time switch over them, api_traits.hpp:91-125). Here only DIM_PROD — the element count of a member array —
 *     is a template parameter; row and plane pitch are run-time values of the hood, so a grid of any shape costs at
 *     most 1.5 x its padded size instead of the enclosing cube. The kernel is instantiated for member strides
 *     2^k and 3 * 2^(k-1), k = B200GEO_SOA_STRIDE_MIN_LOG2 .. B200GEO_SOA_STRIDE_MAX_LOG2 (define them before
 *     including this header to trade compile time against range).
 *   - thread = cell: updateLineX is called for a line of length one, x-adjacent threads touch x-adjacent elements
 *     of every member array (coalesced); `hood[FixedCoord<X, Y, Z>()]` returns a self-contained soa_accessor (no
 *     shared temporary index as in FixedNeighborhood, storage/fixedneighborhood.h:66-83).
 *
 * The SoA member table the engine needs (registration order, element sizes, place in the AoS cell) is derived from
 * the generated accessors themselves (probeMembers below) — the user adds nothing. Requirements: updateLineX must
 * be `__host__ __device__` (as the reference's CUDA path requires of update()), members must be 1, 2, 4 or 8 bytes
 * wide (arrays of those are fine), the cell trivially copyable. Cells with a B200GEO_BIND_CELL line keep their
 * hand-written kernels; SoA cells that also have a per-cell update() take the word-sliced path of b200generic.h
 * (B200Generic::SelectBinding).
 *
 * This header also compiles with a host compiler: everything but the kernel launch is host code, which is how the
 * CPU test suite runs the address arithmetic against the reference (tests/facade/generic_soa_host_test.cpp supplies
 * a host loop as SWEEP over the mock engine's memory; the product ships DeviceSweep only).
 */
#ifndef LIBGEODECOMP_B200_B200GENERICSOA_H
#define LIBGEODECOMP_B200_B200GENERICSOA_H

#include <libflatarray/flat_array.hpp>
#include <libgeodecomp/geometry/fixedcoord.h>
#include <libgeodecomp/misc/apitraits.h>

#include <cstring>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#ifdef __CUDACC__
#include <cuda_runtime.h>
#endif

#ifndef B200GEO_KERNEL_GENERIC
#define B200GEO_KERNEL_GENERIC 0
#endif

/* member strides (elements per member array) the update kernel is instantiated for */
#ifndef B200GEO_SOA_STRIDE_MIN_LOG2
#define B200GEO_SOA_STRIDE_MIN_LOG2 16
#endif
#ifndef B200GEO_SOA_STRIDE_MAX_LOG2
#define B200GEO_SOA_STRIDE_MAX_LOG2 31
#endif

namespace LibGeoDecomp {

namespace B200Generic {

namespace SoA {

static_assert(B200GEO_SOA_STRIDE_MIN_LOG2 >= 9 && B200GEO_SOA_STRIDE_MAX_LOG2 <= 31 &&
              B200GEO_SOA_STRIDE_MIN_LOG2 <= B200GEO_SOA_STRIDE_MAX_LOG2,
              "member strides must be multiples of 256 elements and fit the int index of updateLineX()");

/* entry I of the stride list: 2^MIN, 3 * 2^(MIN - 1), 2^(MIN + 1), ... , 2^MAX */
template<int I>
struct Stride {
    static const long VALUE = (I % 2 == 0) ?
        (1L << (B200GEO_SOA_STRIDE_MIN_LOG2 + I / 2)) :
        (3L << (B200GEO_SOA_STRIDE_MIN_LOG2 + I / 2 - 1));
};

static const int STRIDES = 2 * (B200GEO_SOA_STRIDE_MAX_LOG2 - B200GEO_SOA_STRIDE_MIN_LOG2) + 1;

/* smallest listed stride that holds `least` elements; LibFlatArray's error for grids beyond its size list
 * (macros.hpp:986) */
inline long chooseStride(long least)
{
    for (int i = 0; i < STRIDES; ++i) {
        long v = (i % 2 == 0) ? (1L << (B200GEO_SOA_STRIDE_MIN_LOG2 + i / 2)) : (3L << (B200GEO_SOA_STRIDE_MIN_LOG2 + i / 2 - 1));
        if (v >= least) {
            return v;
        }
    }
    throw std::out_of_range("grid dimension too large");
}

/* The neighbourhood of the cell at element index `index` of the OLD grid: what FixedNeighborhood
 * (storage/fixedneighborhood.h:40-118) is on the CPU. Edge cells and periodic images are real cells of the grid's
 * ghost ring, so no boundary offsets are needed. */
template<typename CELL, long STRIDE>
class Hood
{
public:
    typedef LibFlatArray::soa_accessor<CELL, STRIDE, 1, 1, 0> Accessor;
    typedef CELL Cell;

    __host__ __device__
    Hood(char *data, long index, long pitch, long plane) :
        data(data),
        myIndex(index),
        pitch(pitch),
        plane(plane)
    {}

    template<int X, int Y, int Z>
    __host__ __device__
    inline const Accessor operator[](FixedCoord<X, Y, Z>) const
    {
        return Accessor(data, myIndex + X + Y * pitch + Z * plane);
    }

    __host__ __device__
    inline void operator>>(CELL& cell) const
    {
        Accessor(data, myIndex) >> cell;
    }

    __host__ __device__
    inline long& index()
    {
        return myIndex;
    }

    __host__ __device__
    inline const long& index() const
    {
        return myIndex;
    }

    __host__ __device__
    inline void operator+=(const long offset)
    {
        myIndex += offset;
    }

    __host__ __device__
    inline void operator++()
    {
        ++myIndex;
    }

private:
    char *data;
    long myIndex;
    long pitch;
    long plane;
};

/* the update of ONE cell = a line of length one (updateLineX's loop runs once) */
template<typename CELL, long STRIDE>
__host__ __device__
inline void updateCell(char *oldData, char *newData, long index, long pitch, long plane, unsigned nanoStep)
{
    Hood<CELL, STRIDE> hoodOld(oldData, index, pitch, plane);
    LibFlatArray::soa_accessor<CELL, STRIDE, 1, 1, 0> hoodNew(newData, index);
    CELL::updateLineX(hoodOld, static_cast<int>(index + 1), hoodNew, nanoStep);
}

/* what one sweep over a box needs to know */
struct BoxArgs {
    char *oldData;      /* b200geo_grid_member_ptr(g, 0, 0): the accessors' data pointer of the current buffer */
    char *newData;      /* ... of the scratch buffer */
    long first;         /* element index of the box's first cell */
    long pitch, plane;  /* elements per row / per plane */
    int dim[3];
    unsigned nanoStep;
    void *stream;
};

#ifdef __CUDACC__
template<typename CELL, long STRIDE>
__global__ void __launch_bounds__(256)
updateLineXKernel(char *oldData, char *newData, long first, long pitch, long plane, int nx, int ny, unsigned nanoStep)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int z = blockIdx.z;
    if (x >= nx || y >= ny) {
        return;
    }
    updateCell<CELL, STRIDE>(oldData, newData, first + x + y * pitch + z * plane, pitch, plane, nanoStep);
}

/* the product's SWEEP: one kernel launch per box */
struct DeviceSweep {
    template<typename CELL, long STRIDE>
    static void run(const BoxArgs& a)
    {
        dim3 block(128, a.dim[1] > 1 ? 2 : 1, 1);
        dim3 grid((a.dim[0] + block.x - 1) / block.x, (a.dim[1] + block.y - 1) / block.y, a.dim[2]);
        if (grid.y > 65535u || grid.z > 65535u) {
            throw std::out_of_range("grid dimension too large");
        }
        updateLineXKernel<CELL, STRIDE><<<grid, block, 0, static_cast<cudaStream_t>(a.stream)>>>(
            a.oldData, a.newData, a.first, a.pitch, a.plane, a.dim[0], a.dim[1], a.nanoStep);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) {
            throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(e) + " in generic updateLineX kernel");
        }
    }

    static bool selectDevice(int device)
    {
        return cudaSetDevice(device) == cudaSuccess;
    }
};
#endif

/* run-time member stride -> the instantiation for it (the role of LibFlatArray's size switch,
 * soa_grid.hpp callback / detail/dual_callback_helper.hpp) */
template<typename CELL, typename SWEEP, int I = 0, bool END = (I >= STRIDES)>
struct Dispatch {
    static void run(long stride, const BoxArgs& args)
    {
        if (stride == Stride<I>::VALUE) {
            SWEEP::template run<CELL, Stride<I>::VALUE>(args);
        } else {
            Dispatch<CELL, SWEEP, I + 1>::run(stride, args);
        }
    }
};

template<typename CELL, typename SWEEP, int I>
struct Dispatch<CELL, SWEEP, I, true> {
    static void run(long, const BoxArgs&)
    {
        throw std::logic_error("generic SoA path: no kernel instantiated for this member stride");
    }
};

/* The SoA member table (registration order; array members element by element): for each member element its
 * width and its byte offset inside the AoS cell. LibFlatArray offers no iteration over registered members, but
 * its generated accessors copy a cell member by member (operator<<, macros.hpp:227-235): a cell with ONE marked
 * byte is pushed through soa_accessor<CELL, 2, 1, 1, 0> (two cells per member array, so that the arrays of
 * different members cannot overlap) at element index 0 and 1 — where the mark lands gives the member's place in
 * the aggregated cell, how far it moves between the two gives the element width. */
template<typename CELL>
inline std::vector<B200Member> probeMembers()
{
    static_assert(std::is_trivially_copyable<CELL>::value,
                  "generic SoA path: the cell must be trivially copyable; bind a hand-written kernel otherwise");
    typedef LibFlatArray::soa_accessor<CELL, 2, 1, 1, 0> Accessor;
    const std::size_t aggregated = LibFlatArray::aggregated_member_size<CELL>::VALUE;
    std::vector<long> source(2 * aggregated, -1), width(2 * aggregated, 0);
    std::vector<char> buf(2 * aggregated + 16);
    const char *failure = "generic SoA path: cannot derive the member table of this cell";
    for (std::size_t b = 0; b < sizeof(CELL); ++b) {
        CELL cell = CELL();
        std::memset(reinterpret_cast<char*>(&cell), 0, sizeof(CELL));
        reinterpret_cast<char*>(&cell)[b] = 1;
        long landed[2] = {-1, -1};
        for (int index = 0; index < 2; ++index) {
            std::fill(buf.begin(), buf.end(), 0);
            Accessor accessor(buf.data(), index);
            accessor << cell;
            for (std::size_t p = 0; p < buf.size(); ++p) {
                if (buf[p] != 0) {
                    landed[index] = (long)p;
                    break;
                }
            }
        }
        if (landed[0] < 0) {
            continue;   /* padding, or a member that is not registered */
        }
        if (landed[1] <= landed[0] || (std::size_t)landed[0] >= 2 * aggregated) {
            throw std::logic_error(failure);
        }
        source[landed[0]] = (long)b;
        width[landed[0]] = landed[1] - landed[0];
    }
    /* element e of the aggregated cell (bytes [at, at + w)) occupies bytes [2 * at, 2 * at + w) of the probe buffer */
    std::vector<B200Member> ret;
    for (std::size_t at = 0; at < aggregated;) {
        const std::size_t p = 2 * at;
        if (source[p] < 0 || width[p] <= 0 || at + width[p] > aggregated) {
            throw std::logic_error(failure);
        }
        for (long k = 1; k < width[p]; ++k) {
            if (source[p + k] != source[p] + k || width[p + k] != width[p]) {
                throw std::logic_error(failure);
            }
        }
        B200Member member;
        member.offsetInCell = (std::size_t)source[p];
        member.bytes = (int)width[p];
        ret.push_back(member);
        at += width[p];
    }
    return ret;
}

inline void check(int rc)
{
    if (rc >= 0) {
        return;
    }
    std::string msg = b200geo_last_error();
    switch (rc) {
    case B200GEO_ERR_INVALID:
        throw std::invalid_argument(msg);
    case B200GEO_ERR_LOGIC:
        throw std::logic_error(msg);
    case B200GEO_ERR_OUT_OF_RANGE:
        throw std::out_of_range(msg);
    default:
        throw std::runtime_error(msg.find("CUDA error") == 0 ? msg : "CUDA error: " + msg);
    }
}

/* What B200KernelBinding<CELL> is for an unbound SoA cell. SWEEP enqueues one sweep over a box for a given
 * member stride (DeviceSweep in the product). */
template<typename CELL, typename SWEEP>
struct Binding {
    typedef typename APITraits::SelectTopology<CELL>::Value Topology;
    static const int DIM = Topology::DIM;
    static const unsigned NANO_STEPS = APITraits::SelectNanoSteps<CELL>::VALUE;

    static int kernel()
    {
        return B200GEO_KERNEL_GENERIC;
    }

    static std::vector<B200Member> members()
    {
        static const std::vector<B200Member> table = probeMembers<CELL>();
        return table;
    }

    /* B200Grid<CELL>::create() asks the binding for the grid: uniform element layout, listed member stride */
    static int createGrid(const b200geo_grid_desc *desc, int device, b200geo_grid **out)
    {
        int64_t least = 0;
        int rc = b200geo_grid_uniform_min_stride(desc, &least);
        if (rc < 0) {
            return rc;
        }
        return b200geo_grid_create_uniform(desc, device, chooseStride((long)least), out);
    }

    /* enqueue one sweep over the box (origin, dim) of grid g: current buffer -> scratch buffer */
    static void launchBox(b200geo_grid *g, unsigned nanoStep, const int32_t origin[3], const int32_t dim[3], void *stream)
    {
        if (dim[0] <= 0 || dim[1] <= 0 || dim[2] <= 0) {
            return;
        }
        int64_t first = 0, pitch = 0, plane = 0, stride = 0;
        void *oldData = 0, *newData = 0;
        check(b200geo_grid_layout(g, 0, &pitch, &plane, &first));
        check(b200geo_grid_member_stride(g, &stride));
        check(b200geo_grid_member_ptr(g, 0, 0, &oldData));
        check(b200geo_grid_member_ptr(g, 0, 1, &newData));
        if (stride <= 0) {
            throw std::logic_error("generic SoA path: the grid is not in the uniform element layout");
        }
        BoxArgs args;
        args.oldData = static_cast<char*>(oldData);
        args.newData = static_cast<char*>(newData);
        args.first = (long)(first + origin[0] + origin[1] * pitch + origin[2] * plane);
        args.pitch = (long)pitch;
        args.plane = (long)plane;
        for (int i = 0; i < 3; ++i) {
            args.dim[i] = dim[i];
        }
        args.nanoStep = nanoStep % NANO_STEPS;
        args.stream = stream;
        Dispatch<CELL, SWEEP>::run((long)stride, args);
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
            check(b200geo_refresh_ghosts(g, 0));
            launchBox(g, firstNanoStep + t, origin, dim, 0);
            check(b200geo_swap(g));
        }
    }

    /* on a slab group (B200StripingSimulator) the group drives the schedule — rims, halo copies, interiors — and
     * calls back for every box it wants updated */
    static int updateCallback(void *, b200geo_grid *g, uint32_t nanoStep, const int32_t origin[3], const int32_t dim[3], void *stream)
    {
        try {
            int device = 0;
            check(b200geo_grid_device(g, &device));
            if (!SWEEP::selectDevice(device)) {
                return B200GEO_ERR_CUDA;
            }
            launchBox(g, nanoStep, origin, dim, stream);
        } catch (const std::exception&) {
            return B200GEO_ERR_CUDA;
        }
        return B200GEO_OK;
    }

    static void groupStep(b200geo_group *group, unsigned firstNanoStep, unsigned sweeps)
    {
        check(b200geo_group_step_with(group, &updateCallback, 0, firstNanoStep, sweeps));
    }
};

}

}

}

#endif
