/* Checkpoints in the reference's MPI-IO file layout (include/libgeodecomp_b200/b200checkpoint.h):
 *  - the bytes B200CheckpointWriter puts on disk from the DEVICE grid equal (a) the layout io/mpiio.h:82-130,195-203
 *    defines — header {Coord<DIM>, step, maxSteps, edge cell}, then the cells row-major — built here by hand from the
 *    reference SerialSimulator's grid, and (b) what the same writer produces on the reference's host grid;
 *  - B200ParallelCheckpointWriter on B200StripingSimulator (every slab writes its own region into the one file)
 *    gives the same file;
 *  - B200CheckpointInitializer restarts a run from such a file (metadata from the header, cells box by box) on
 *    B200Simulator, B200StripingSimulator and on the reference's SerialSimulator, and the restarted runs end where
 *    the uninterrupted one does;
 *  - single-member cells (Jacobi, 3-D; Game of Life, 2-D) and a 24-member cell (LBM: members gathered / un-sliced),
 *    transfer pieces smaller than the grid (ragged pieces).
 * The reference's own MPIIOWriter / MPIIOInitializer need MPI and cannot be built here (no MPI in the image).
 * Linked against libb200geo.so (GPU box) or tests/facade/mock_b200geo.cpp (CPU suite). */
#include <libgeodecomp_b200/b200checkpoint.h>
#include <libgeodecomp_b200/b200stripingsimulator.h>

#include "fixtures.h"

#include <unistd.h>

static std::vector<char> slurp(const std::string& name)
{
    std::ifstream in(name.c_str(), std::ios::binary);
    return std::vector<char>((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
}

/* `slabs` slabs, round-robin over the devices present */
static std::vector<int> devicesFor(int slabs)
{
    int n = b200geo_device_count();
    std::vector<int> ret;
    for (int i = 0; i < slabs; ++i) {
        ret.push_back(n > 0 ? i % n : 0);
    }
    return ret;
}

static std::string scratch(const char *tag)
{
    char buf[256];
    std::snprintf(buf, sizeof(buf), "/tmp/b200geo_ckpt_%d_%s_", (int)getpid(), tag);
    return buf;
}

/* the file MPIIO::writeRegion would write for the whole grid (io/mpiio.h:96-127) */
template<typename CELL, int DIM>
static std::vector<char> expectedFile(const GridBase<CELL, DIM>& grid, unsigned step, unsigned maxSteps)
{
    Coord<DIM> dim = grid.dimensions();
    std::vector<char> out(sizeof(int) * DIM + 2 * sizeof(unsigned) + sizeof(CELL) + (std::size_t)dim.prod() * sizeof(CELL));
    std::size_t at = 0;
    for (int d = 0; d < DIM; ++d) {
        int v = dim[d];
        std::memcpy(&out[at], &v, sizeof(int));
        at += sizeof(int);
    }
    std::memcpy(&out[at], &step, sizeof(unsigned));
    at += sizeof(unsigned);
    std::memcpy(&out[at], &maxSteps, sizeof(unsigned));
    at += sizeof(unsigned);
    CELL edge = grid.getEdge();
    std::memcpy(&out[at], &edge, sizeof(CELL));
    at += sizeof(CELL);
    CoordBox<DIM> box = grid.boundingBox();
    for (typename CoordBox<DIM>::Iterator i = box.begin(); i != box.end(); ++i) {
        CELL c = grid.get(*i);
        std::memcpy(&out[at + (std::size_t)(*i - box.origin).toIndex(dim) * sizeof(CELL)], &c, sizeof(CELL));
    }
    return out;
}

template<typename CELL, int DIM>
static long differingCells(const GridBase<CELL, DIM> *a, const GridBase<CELL, DIM> *b)
{
    long bad = 0;
    CoordBox<DIM> box = a->boundingBox();
    for (typename CoordBox<DIM>::Iterator i = box.begin(); i != box.end(); ++i) {
        bad += !(a->get(*i) == b->get(*i));
    }
    return bad;
}

template<typename CELL, typename INIT, int DIM>
static void checkpointAndRestart(const char *name, const Coord<DIM>& dim, unsigned steps, unsigned period, std::size_t pieceBytes)
{
    const std::string onDevice = scratch("dev"), onHost = scratch("host"), onSlabs = scratch("slabs");
    /* the uninterrupted runs, checkpointing every `period` steps */
    SerialSimulator<CELL> ref(new INIT(dim, steps));
    ref.addWriter(new B200CheckpointWriter<CELL>(onHost, period, steps, pieceBytes));
    B200Simulator<CELL> dev(new INIT(dim, steps));
    dev.addWriter(new B200CheckpointWriter<CELL>(onDevice, period, steps, pieceBytes));
    B200StripingSimulator<CELL> slabs(new INIT(dim, steps), devicesFor(3));
    slabs.addWriter(new B200ParallelCheckpointWriter<CELL>(onSlabs, period, steps, pieceBytes));
    ref.run();
    dev.run();
    slabs.run();

    /* the layout, at the final step, against a hand-built file */
    std::vector<char> want = expectedFile<CELL, DIM>(*ref.getGrid(), steps, steps);
    bool layout = slurp(B200CheckpointHelpers::filename(onDevice, steps)) == want;
    CHECK(layout);
    /* every checkpoint: device == host grid == slabs */
    bool same = true;
    unsigned files = 0;
    for (unsigned s = 0; s <= steps; s += period) {
        std::vector<char> a = slurp(B200CheckpointHelpers::filename(onHost, s));
        same &= !a.empty() && a == slurp(B200CheckpointHelpers::filename(onDevice, s)) && a == slurp(B200CheckpointHelpers::filename(onSlabs, s));
        ++files;
    }
    CHECK(same);

    /* restart from the checkpoint in the middle */
    const unsigned mid = (steps / period / 2) * period > 0 ? (steps / period / 2) * period : period;
    const std::string file = B200CheckpointHelpers::filename(onDevice, mid);
    B200CheckpointInitializer<CELL> probe(file);
    CHECK(probe.startStep() == mid && probe.maxSteps() == steps && probe.gridDimensions() == dim);
    B200Simulator<CELL> again(new B200CheckpointInitializer<CELL>(file, pieceBytes));
    B200StripingSimulator<CELL> againSlabs(new B200CheckpointInitializer<CELL>(file, pieceBytes), devicesFor(2));
    SerialSimulator<CELL> againRef(new B200CheckpointInitializer<CELL>(file, pieceBytes));
    CHECK(again.getStep() == mid);
    again.run();
    againSlabs.run();
    againRef.run();
    CHECK(again.getStep() == steps && againRef.getStep() == steps);
    long bad = differingCells<CELL, DIM>(ref.getGrid(), again.getGrid()) + differingCells<CELL, DIM>(ref.getGrid(), againSlabs.getGrid()) +
               differingCells<CELL, DIM>(ref.getGrid(), againRef.getGrid());
    CHECK(bad == 0);
    CHECK(again.getGrid()->getEdge() == ref.getGrid()->getEdge());

    std::printf("%-12s %u checkpoints of %zu bytes: layout %s, device / host grid / 3 slabs %s; restart at step %u of %u: %ld cells differ\n",
                name, files, want.size(), layout ? "as io/mpiio.h" : "DIFFERENT", same ? "identical" : "DIFFERENT", mid, steps, bad);
    for (unsigned s = 0; s <= steps; s += period) {
        std::remove(B200CheckpointHelpers::filename(onHost, s).c_str());
        std::remove(B200CheckpointHelpers::filename(onDevice, s).c_str());
        std::remove(B200CheckpointHelpers::filename(onSlabs, s).c_str());
    }
}

int main()
{
    try {
        checkpointAndRestart<Jacobi7Cube, SeededInitializer<Jacobi7Cube>, 3>("Jacobi7Cube", Coord<3>(23, 9, 12), 8, 2, 64 << 20);
        checkpointAndRestart<Jacobi7Cube, SeededInitializer<Jacobi7Cube>, 3>("Jacobi7Cube", Coord<3>(23, 9, 12), 6, 3, 1000);   /* ragged pieces */
        checkpointAndRestart<Jacobi27Torus, SeededInitializer<Jacobi27Torus>, 3>("Jacobi27Torus", Coord<3>(12, 6, 8), 4, 2, 4096);
        checkpointAndRestart<ConwayCube, SeededInitializer<ConwayCube>, 2>("ConwayCube", Coord<2>(70, 33), 12, 4, 500);
        checkpointAndRestart<LBMCellF, LBMInitializer, 3>("LBMCellF", Coord<3>(14, 9, 8), 6, 3, 20000);
        bool missing = false;
        try {
            B200CheckpointInitializer<Jacobi7Cube> none("/nonexistent/b200geo.mpiio");
        } catch (const std::runtime_error&) {
            missing = true;
        }
        CHECK(missing);
    } catch (const std::exception& e) {
        std::printf("FAILED with exception: %s\n", e.what());
        return 2;
    }
    if (failures) {
        std::printf("%d check(s) FAILED\n", failures);
        return 1;
    }
    std::printf("checkpoint_test: all checks passed\n");
    return 0;
}
