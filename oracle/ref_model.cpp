/* TEST INFRASTRUCTURE — one translation unit per oracle binary; the model is selected with
 * -DB200_MODEL_<NAME> by oracle/Makefile (SoA cells cost ~1 min of template instantiation
 * each, so they are built as separate binaries in parallel). See ref_driver.h. */
#include "ref_driver.h"

#if defined(B200_MODEL_JACOBI6CUBE) || defined(B200_MODEL_JACOBI6TORUS) || \
    defined(B200_MODEL_JACOBI7CUBE) || defined(B200_MODEL_JACOBI7TORUS) || \
    defined(B200_MODEL_JACOBI27CUBE) || defined(B200_MODEL_JACOBI27TORUS)
#include "models/jacobi.h"
#define B200_JACOBI_FAMILY
#endif
#if defined(B200_MODEL_CONWAYCUBE) || defined(B200_MODEL_CONWAYTORUS)
#include "models/conway.h"
#endif
#if defined(B200_MODEL_LBM)
#include "models/lbm.h"
#endif
#if defined(B200_MODEL_NBODY)
#include "models/nbody.h"
#endif
#if defined(B200_MODEL_CONTAINER)
#include "models/container.h"
#endif

#if defined(B200_MODEL_JACOBI6CUBE)
typedef b200models::Jacobi6Cube Model;
#elif defined(B200_MODEL_JACOBI6TORUS)
typedef b200models::Jacobi6Torus Model;
#elif defined(B200_MODEL_JACOBI7CUBE)
typedef b200models::Jacobi7Cube Model;
#elif defined(B200_MODEL_JACOBI7TORUS)
typedef b200models::Jacobi7Torus Model;
#elif defined(B200_MODEL_JACOBI27CUBE)
typedef b200models::Jacobi27Cube Model;
#elif defined(B200_MODEL_JACOBI27TORUS)
typedef b200models::Jacobi27Torus Model;
#elif defined(B200_MODEL_CONWAYCUBE)
typedef b200models::ConwayCube Model;
#elif defined(B200_MODEL_CONWAYTORUS)
typedef b200models::ConwayTorus Model;
#elif defined(B200_MODEL_LBM)
typedef b200models::LBMCellF Model;
#elif defined(B200_MODEL_NBODY)
typedef b200models::NBodyCell Model;
#elif defined(B200_MODEL_CONTAINER)
typedef LibGeoDecomp::ContainerCell<b200models::MeshElement<3, false>, b200models::CONTAINER_CAPACITY> Model;
#else
#error "select a model with -DB200_MODEL_<NAME>"
#endif

namespace refdriver {

#ifdef B200_JACOBI_FAMILY
template<> struct Codec<Model> {
    static const int BYTES = 8;
    static Model edge(double v) { return Model(v); }
    static void fromRaw(Model *c, const char *raw, std::size_t cells, std::size_t i)
    { c->temp = rawGet<double>(raw, cells, 0, i); }
    static void toRaw(const Model& c, char *raw, std::size_t cells, std::size_t i)
    { rawPut<double>(raw, cells, 0, i, c.temp); }
};
#endif

#if defined(B200_MODEL_CONWAYCUBE) || defined(B200_MODEL_CONWAYTORUS)
template<> struct Codec<Model> {
    static const int BYTES = 1;
    static Model edge(double v) { return Model(v != 0); }
    static void fromRaw(Model *c, const char *raw, std::size_t cells, std::size_t i)
    { c->alive = rawGet<unsigned char>(raw, cells, 0, i) != 0; }
    static void toRaw(const Model& c, char *raw, std::size_t cells, std::size_t i)
    { rawPut<unsigned char>(raw, cells, 0, i, c.alive ? 1 : 0); }
};
#endif

#if defined(B200_MODEL_LBM)
template<> struct Codec<Model> {
    static const int BYTES = 24 * 4;
    static Model edge(double v) { return Model((float)v); }
    static float *member(Model *c, int m)
    {
        float *tab[23] = {
            &c->C, &c->N, &c->E, &c->W, &c->S, &c->T, &c->B, &c->NW, &c->SW, &c->NE, &c->SE,
            &c->TW, &c->BW, &c->TE, &c->BE, &c->TN, &c->BN, &c->TS, &c->BS,
            &c->density, &c->velocityX, &c->velocityY, &c->velocityZ };
        return tab[m];
    }
    static void fromRaw(Model *c, const char *raw, std::size_t cells, std::size_t i)
    {
        for (int m = 0; m < 23; ++m) *member(c, m) = rawGet<float>(raw, cells, 4 * m, i);
        c->state = rawGet<int>(raw, cells, 4 * 23, i);
    }
    static void toRaw(const Model& c, char *raw, std::size_t cells, std::size_t i)
    {
        for (int m = 0; m < 23; ++m) rawPut<float>(raw, cells, 4 * m, i, *member(const_cast<Model*>(&c), m));
        rawPut<int>(raw, cells, 4 * 23, i, c.state);
    }
};
#endif

}

#if defined(B200_MODEL_NBODY)
int main(int argc, char **argv) { return b200models::nbodyMain(argc, argv); }
#elif defined(B200_MODEL_CONTAINER)
int main(int argc, char **argv) { return b200models::containerMain(argc, argv); }
#else
int main(int argc, char **argv)
{
    try {
        return refdriver::runModel<Model>(argc > 1 ? argv[1] : "?", argc, argv);
    } catch (const std::exception& e) {
        fprintf(stderr, "error: %s\n", e.what());
        return 4;
    }
}
#endif
