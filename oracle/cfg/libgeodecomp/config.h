/* Hand-written stand-in for the header the reference's CMake normally generates
 * (CMakeModules/util.cmake:7-24, lgd_dump_config). Test infrastructure only. */
#ifndef LIBGEODECOMP_CONFIG_H
#define LIBGEODECOMP_CONFIG_H
#define LIBGEODECOMP_DEBUG_LEVEL 0
#define LIBGEODECOMP_WITH_CPP14 true
#ifdef _OPENMP
#define LIBGEODECOMP_WITH_THREADS true
#endif
#endif
