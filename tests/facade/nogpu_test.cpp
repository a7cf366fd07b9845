/* Host-only check of the C++ façade on a machine WITHOUT a usable CUDA device: there is no CPU fallback —
 * constructing a device grid or a simulator fails loudly with the reference's exception for CUDA failures
 * (std::runtime_error("CUDA error ..."), misc/cudautil.h:48-55), and argument validation still maps to
 * std::invalid_argument / std::logic_error before any CUDA call. Run by tests/test_capi_symbols.py when no GPU
 * is present; on a GPU box it only checks the validation part. */
#include "fixtures.h"

#include <libgeodecomp_b200/b200stripingsimulator.h>

int main()
{
    int devices = b200geo_device_count();
    std::printf("nogpu_test: b200geo_device_count() = %d\n", devices);
    if (devices <= 0) {
        bool loud = false;
        try {
            B200Grid<Jacobi7Cube> grid(CoordBox<3>(Coord<3>(), Coord<3>(8, 8, 8)));
        } catch (const std::runtime_error& e) {
            loud = std::string(e.what()).find("CUDA error") == 0;
            std::printf("B200Grid without a device: std::runtime_error(\"%s\")\n", e.what());
        }
        CHECK(loud);
        loud = false;
        try {
            B200Simulator<Jacobi7Cube> sim(new SeededInitializer<Jacobi7Cube>(Coord<3>(8, 8, 8), 2));
            sim.run();
        } catch (const std::runtime_error& e) {
            loud = std::string(e.what()).find("CUDA error") == 0;
        }
        CHECK(loud);
        loud = false;
        try {
            B200StripingSimulator<Jacobi7Cube> sim(new SeededInitializer<Jacobi7Cube>(Coord<3>(8, 8, 8), 2));
            sim.run();
        } catch (const std::runtime_error& e) {
            loud = std::string(e.what()).find("CUDA error") == 0;
        }
        CHECK(loud);
    }
    // validation happens before any CUDA call
    b200geo_grid_desc desc;
    std::memset(&desc, 0, sizeof(desc));
    b200geo_grid *g = 0;
    desc.dim[0] = desc.dim[1] = desc.dim[2] = 4;
    desc.n_members = 1;
    desc.member_bytes[0] = 3;
    bool invalid = false;
    try {
        B200Helpers::check(b200geo_grid_create(&desc, 0, &g));
    } catch (const std::invalid_argument&) {
        invalid = true;
    }
    CHECK(invalid);
    desc.member_bytes[0] = 8;
    desc.ghost[0] = desc.ghost[1] = desc.ghost[2] = 1;
    desc.ghost_mode[0][0] = B200GEO_GHOST_PEER;     // PEER layers exist on the slab axis only
    bool logic = false;
    try {
        B200Helpers::check(b200geo_grid_create(&desc, 0, &g));
    } catch (const std::logic_error&) {
        logic = true;
    }
    CHECK(logic);
    bool thin = false;
    try {
        B200StripedGrid<Jacobi7Cube> grid(CoordBox<3>(Coord<3>(), Coord<3>(8, 8, 4)), std::vector<int>(4, 0), 2);
    } catch (const std::invalid_argument&) {
        thin = true;
    } catch (const std::runtime_error&) {
        thin = devices <= 0;   // without a device the first slab already fails
    }
    CHECK(thin);
    if (failures) {
        std::printf("%d check(s) FAILED\n", failures);
        return 1;
    }
    std::printf("nogpu_test: all checks passed\n");
    return 0;
}
