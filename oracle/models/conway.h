/* TEST INFRASTRUCTURE — Conway's Game of Life as a LibGeoDecomp user model; same rule and
 * same run-time Coord<2> neighbourhood access as src/examples/gameoflife/main.cpp:25-62
 * (vanilla update path: storage/vanillaupdatefunctor.h:12-36, default Moore<2,1> stencil and
 * Cube<2> topology, misc/apitraits.h:267,328). The Torus variant only adds the topology trait.
 */
#ifndef B200GEO_ORACLE_MODELS_CONWAY_H
#define B200GEO_ORACLE_MODELS_CONWAY_H

#include <libgeodecomp/misc/apitraits.h>
#include <libgeodecomp/geometry/coord.h>

namespace b200models {

using namespace LibGeoDecomp;

#define B200_DEFINE_CONWAY(NAME, API_BASES)                                     \
class NAME                                                                      \
{                                                                               \
public:                                                                         \
    class API API_BASES                                                         \
    {};                                                                         \
                                                                                \
    explicit NAME(bool alive = false) : alive(alive) {}                         \
                                                                                \
    template<typename HOOD>                                                     \
    void update(const HOOD& hood, unsigned /* nanoStep */)                      \
    {                                                                           \
        int living = 0;                                                         \
        for (int y = -1; y < 2; ++y) {                                          \
            for (int x = -1; x < 2; ++x) {                                      \
                living += hood[Coord<2>(x, y)].alive;                           \
            }                                                                   \
        }                                                                       \
        bool self = hood[Coord<2>(0, 0)].alive;                                 \
        living -= self;                                                         \
        alive = self ? ((2 <= living) && (living <= 3)) : (living == 3);        \
    }                                                                           \
                                                                                \
    bool operator==(const NAME& o) const { return alive == o.alive; }           \
                                                                                \
    bool alive;                                                                 \
};

#define B200_NO_BASES
#define B200_TORUS2_BASES : public APITraits::HasTorusTopology<2>
B200_DEFINE_CONWAY(ConwayCube,  B200_NO_BASES)
B200_DEFINE_CONWAY(ConwayTorus, B200_TORUS2_BASES)

}

#endif
