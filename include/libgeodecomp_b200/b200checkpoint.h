/* Checkpoints in the reference's MPI-IO file layout, written and read without MPI and in whole boxes.
 *
 * File layout = what MPIIO<CELL>::writeRegion produces (io/mpiio.h:82-130, getLengths :195-203):
 *     Coord<DIM> dimensions | unsigned step | unsigned maxSteps | CELL edgeCell | CELL cells[prod(dimensions)]
 * cells row-major (x fastest: offset = headerLength + c.toIndex(dimensions) * sizeof(CELL), mpiio.h:183-190), in
 * the cell's in-memory (AoS) form — the extent Typemaps generates for a cell struct is sizeof(CELL). Files named
 * <prefix><step, 5 digits>.mpiio (io/mpiiowriter.h:66-71), so a run checkpointed here restarts under the
 * reference's MPIIOInitializer and vice versa.
 *
 *   B200CheckpointWriter<CELL>          Writer (io/mpiiowriter.h:20-72) for B200Simulator / SerialSimulator
 *   B200ParallelCheckpointWriter<CELL>  ParallelWriter (io/parallelmpiiowriter.h) for B200StripingSimulator: every
 *                                       call writes its validRegion at the cells' places in the one file
 *                                       (pwrite: disjoint regions of several processes do not need MPI-IO for that)
 *   B200CheckpointInitializer<CELL>     Initializer (io/mpiioinitializer.h:22-78): metadata from the header,
 *                                       grid() fills target->boundingBox()
 *
 * The reference moves a checkpoint streak by streak (grid.get(streak) / MPI_File_write per streak; on its CUDA
 * grids that is one cudaMemcpy per streak, storage/cudagrid.h:181-202). Here the region is cut into boxes of whole
 * planes / rows of about 64 MiB, and each box crosses the bus in ONE transfer (GridBase::saveRegion / loadRegion with
 * an AoS buffer — B200Grid gathers the members on the device and un-slices them once per box) and reaches the file in
 * one pwrite per contiguous run of rows. Grids that offer only char buffers (SoAGrid) fall back to streaks.
 */
#ifndef LIBGEODECOMP_B200_B200CHECKPOINT_H
#define LIBGEODECOMP_B200_B200CHECKPOINT_H

#include <libgeodecomp/io/initializer.h>
#include <libgeodecomp/io/parallelwriter.h>
#include <libgeodecomp/io/writer.h>
#include <libgeodecomp/misc/clonable.h>

#include <fcntl.h>
#include <sys/stat.h>
#include <sys/types.h>
#include <unistd.h>

#include <cerrno>
#include <cstring>
#include <iomanip>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace LibGeoDecomp {

namespace B200CheckpointHelpers {

template<typename CELL, int DIM>
struct Layout {
    static std::size_t headerLength()
    {
        return sizeof(int) * DIM + 2 * sizeof(unsigned) + sizeof(CELL);
    }

    static off_t offset(const Coord<DIM>& c, const Coord<DIM>& dimensions)
    {
        return (off_t)headerLength() + (off_t)c.toIndex(dimensions) * (off_t)sizeof(CELL);
    }
};

inline std::string filename(const std::string& prefix, unsigned step)
{
    std::ostringstream buf;
    buf << prefix << std::setfill('0') << std::setw(5) << step << ".mpiio";
    return buf.str();
}

class File
{
public:
    File(const std::string& name, bool write) :
        fd(::open(name.c_str(), write ? (O_CREAT | O_WRONLY) : O_RDONLY, 0644)),
        name(name)
    {
        if (fd < 0) {
            throw std::runtime_error("could not open checkpoint file " + name + ": " + std::strerror(errno));
        }
    }

    ~File()
    {
        ::close(fd);
    }

    void write(const void *data, std::size_t bytes, off_t at)
    {
        const char *p = static_cast<const char*>(data);
        while (bytes > 0) {
            ssize_t n = ::pwrite(fd, p, bytes, at);
            if (n <= 0) {
                throw std::runtime_error("write to checkpoint file " + name + " failed: " + std::strerror(errno));
            }
            p += n;
            at += n;
            bytes -= (std::size_t)n;
        }
    }

    void read(void *data, std::size_t bytes, off_t at)
    {
        char *p = static_cast<char*>(data);
        while (bytes > 0) {
            ssize_t n = ::pread(fd, p, bytes, at);
            if (n <= 0) {
                throw std::runtime_error("checkpoint file " + name + " is shorter than its header says");
            }
            p += n;
            at += n;
            bytes -= (std::size_t)n;
        }
    }

private:
    int fd;
    std::string name;
};

/* cut a region into pieces of at most `budget` cells along its streak order (whole streaks; box-shaped regions give
 * box-shaped pieces) */
template<int DIM>
inline std::vector<Region<DIM> > pieces(const Region<DIM>& region, std::size_t budget)
{
    std::vector<Region<DIM> > ret;
    Region<DIM> cur;
    std::size_t cells = 0;
    for (typename Region<DIM>::StreakIterator i = region.beginStreak(); i != region.endStreak(); ++i) {
        if (cells > 0 && cells + i->length() > budget) {
            ret.push_back(cur);
            cur.clear();
            cells = 0;
        }
        cur << *i;
        cells += i->length();
    }
    if (cells > 0) {
        ret.push_back(cur);
    }
    return ret;
}

/* region -> file. Streaks that follow each other in the file (whole rows of the simulation space) become one pwrite. */
template<typename CELL, int DIM, typename GRID>
inline void writeRegion(const GRID& grid, const Coord<DIM>& dimensions, File *file, const Region<DIM>& region, std::size_t budgetCells)
{
    std::vector<Region<DIM> > parts = pieces(region, budgetCells);
    std::vector<CELL> buffer;
    for (std::size_t p = 0; p < parts.size(); ++p) {
        const Region<DIM>& part = parts[p];
        buffer.resize(part.size());
        bool boxTransfer = true;
        try {
            grid.saveRegion(&buffer, part);             /* one transfer per piece */
        } catch (const std::logic_error&) {
            boxTransfer = false;                        /* a grid without AoS buffers: streak by streak */
        }
        if (!boxTransfer) {
            std::size_t at = 0;
            for (typename Region<DIM>::StreakIterator i = part.beginStreak(); i != part.endStreak(); ++i) {
                grid.get(*i, &buffer[at]);
                at += i->length();
            }
        }
        std::size_t at = 0, runStart = 0;
        off_t runOffset = 0, next = -1;
        for (typename Region<DIM>::StreakIterator i = part.beginStreak(); i != part.endStreak(); ++i) {
            off_t here = Layout<CELL, DIM>::offset(i->origin, dimensions);
            if (here != next) {
                if (at > runStart) {
                    file->write(&buffer[runStart], (at - runStart) * sizeof(CELL), runOffset);
                }
                runStart = at;
                runOffset = here;
            }
            at += i->length();
            next = here + (off_t)i->length() * (off_t)sizeof(CELL);
        }
        if (at > runStart) {
            file->write(&buffer[runStart], (at - runStart) * sizeof(CELL), runOffset);
        }
    }
}

template<typename CELL, int DIM>
inline void writeHeader(File *file, const Coord<DIM>& dimensions, unsigned step, unsigned maxSteps, const CELL& edgeCell)
{
    std::vector<char> header(Layout<CELL, DIM>::headerLength());
    std::size_t at = 0;
    for (int d = 0; d < DIM; ++d) {
        int v = dimensions[d];
        std::memcpy(&header[at], &v, sizeof(int));
        at += sizeof(int);
    }
    std::memcpy(&header[at], &step, sizeof(unsigned));
    at += sizeof(unsigned);
    std::memcpy(&header[at], &maxSteps, sizeof(unsigned));
    at += sizeof(unsigned);
    std::memcpy(&header[at], &edgeCell, sizeof(CELL));
    file->write(header.data(), header.size(), 0);
}

}

template<typename CELL_TYPE>
class B200CheckpointWriter : public Clonable<Writer<CELL_TYPE>, B200CheckpointWriter<CELL_TYPE> >
{
public:
    typedef typename Writer<CELL_TYPE>::GridType GridType;
    typedef typename Writer<CELL_TYPE>::Topology Topology;
    static const int DIM = Topology::DIM;
    using Writer<CELL_TYPE>::period;
    using Writer<CELL_TYPE>::prefix;

    /* the constructor of MPIIOWriter without the communicator; transferBytes = size of one device <-> host piece */
    B200CheckpointWriter(const std::string& prefix, const unsigned period, const unsigned maxSteps, std::size_t transferBytes = 64 << 20) :
        Clonable<Writer<CELL_TYPE>, B200CheckpointWriter<CELL_TYPE> >(prefix, period),
        maxSteps(maxSteps),
        budgetCells(transferBytes / sizeof(CELL_TYPE) > 0 ? transferBytes / sizeof(CELL_TYPE) : 1)
    {}

    virtual void stepFinished(const GridType& grid, unsigned step, WriterEvent event)
    {
        if ((event == WRITER_STEP_FINISHED) && (step % period != 0)) {
            return;
        }
        Region<DIM> region;
        region << grid.boundingBox();
        B200CheckpointHelpers::File file(B200CheckpointHelpers::filename(prefix, step), true);
        B200CheckpointHelpers::writeHeader<CELL_TYPE, DIM>(&file, grid.dimensions(), step, maxSteps, grid.getEdge());
        B200CheckpointHelpers::writeRegion<CELL_TYPE, DIM>(grid, grid.dimensions(), &file, region, budgetCells);
    }

private:
    unsigned maxSteps;
    std::size_t budgetCells;
};

template<typename CELL_TYPE>
class B200ParallelCheckpointWriter : public Clonable<ParallelWriter<CELL_TYPE>, B200ParallelCheckpointWriter<CELL_TYPE> >
{
public:
    typedef typename ParallelWriter<CELL_TYPE>::GridType GridType;
    typedef typename APITraits::SelectTopology<CELL_TYPE>::Value Topology;
    static const int DIM = Topology::DIM;
    using ParallelWriter<CELL_TYPE>::period;
    using ParallelWriter<CELL_TYPE>::prefix;

    B200ParallelCheckpointWriter(const std::string& prefix, const unsigned period, const unsigned maxSteps, std::size_t transferBytes = 64 << 20) :
        Clonable<ParallelWriter<CELL_TYPE>, B200ParallelCheckpointWriter<CELL_TYPE> >(prefix, period),
        maxSteps(maxSteps),
        budgetCells(transferBytes / sizeof(CELL_TYPE) > 0 ? transferBytes / sizeof(CELL_TYPE) : 1)
    {}

    virtual void stepFinished(
        const GridType& grid,
        const Region<DIM>& validRegion,
        const Coord<DIM>& globalDimensions,
        unsigned step,
        WriterEvent event,
        std::size_t rank,
        bool /* lastCall */)
    {
        if ((event == WRITER_STEP_FINISHED) && (step % period != 0)) {
            return;
        }
        B200CheckpointHelpers::File file(B200CheckpointHelpers::filename(prefix, step), true);
        if (rank == 0) {
            B200CheckpointHelpers::writeHeader<CELL_TYPE, DIM>(&file, globalDimensions, step, maxSteps, grid.getEdge());
        }
        B200CheckpointHelpers::writeRegion<CELL_TYPE, DIM>(grid, globalDimensions, &file, validRegion, budgetCells);
    }

private:
    unsigned maxSteps;
    std::size_t budgetCells;
};

template<typename CELL_TYPE>
class B200CheckpointInitializer : public Initializer<CELL_TYPE>
{
public:
    typedef typename APITraits::SelectTopology<CELL_TYPE>::Value Topology;
    static const int DIM = Topology::DIM;

    explicit B200CheckpointInitializer(const std::string& filename, std::size_t transferBytes = 64 << 20) :
        file(filename),
        budgetCells(transferBytes / sizeof(CELL_TYPE) > 0 ? transferBytes / sizeof(CELL_TYPE) : 1)
    {
        /* MPIIO::readMetadata, io/mpiio.h:67-80 */
        B200CheckpointHelpers::File in(file, false);
        int dims[DIM];
        in.read(dims, sizeof(dims), 0);
        for (int d = 0; d < DIM; ++d) {
            dimensions[d] = dims[d];
        }
        in.read(&currentStep, sizeof(unsigned), sizeof(dims));
        in.read(&maximumSteps, sizeof(unsigned), sizeof(dims) + sizeof(unsigned));
    }

    /* MPIIO::readRegion (io/mpiio.h:26-65) over target->boundingBox(), piece by piece */
    virtual void grid(GridBase<CELL_TYPE, DIM> *target)
    {
        typedef B200CheckpointHelpers::Layout<CELL_TYPE, DIM> Layout;
        B200CheckpointHelpers::File in(file, false);
        CELL_TYPE edge;
        in.read(&edge, sizeof(CELL_TYPE), (off_t)(Layout::headerLength() - sizeof(CELL_TYPE)));
        target->setEdge(edge);

        Region<DIM> region;
        region << target->boundingBox();
        std::vector<Region<DIM> > parts = B200CheckpointHelpers::pieces(region, budgetCells);
        std::vector<CELL_TYPE> buffer;
        for (std::size_t p = 0; p < parts.size(); ++p) {
            const Region<DIM>& part = parts[p];
            buffer.resize(part.size());
            std::size_t at = 0;
            for (typename Region<DIM>::StreakIterator i = part.beginStreak(); i != part.endStreak(); ++i) {
                /* on Torus topologies the coordinates may lie outside the bounding box */
                Coord<DIM> c = Topology::normalize(i->origin, dimensions);
                in.read(&buffer[at], (std::size_t)i->length() * sizeof(CELL_TYPE), Layout::offset(c, dimensions));
                at += i->length();
            }
            bool boxTransfer = true;
            try {
                target->loadRegion(buffer, part);       /* one transfer per piece */
            } catch (const std::logic_error&) {
                boxTransfer = false;
            }
            if (!boxTransfer) {
                at = 0;
                for (typename Region<DIM>::StreakIterator i = part.beginStreak(); i != part.endStreak(); ++i) {
                    target->set(*i, &buffer[at]);
                    at += i->length();
                }
            }
        }
    }

    virtual Coord<DIM> gridDimensions() const
    {
        return dimensions;
    }

    virtual unsigned maxSteps() const
    {
        return maximumSteps;
    }

    virtual unsigned startStep() const
    {
        return currentStep;
    }

private:
    std::string file;
    std::size_t budgetCells;
    unsigned currentStep;
    unsigned maximumSteps;
    Coord<DIM> dimensions;
};

}

#endif
