/* Hand-written stand-in for the header the reference's CMake normally generates
 * (CMakeModules/util.cmake:7-24, lgd_dump_config). Test infrastructure only. */
#ifndef LIBGEODECOMP_CONFIG_H
#define LIBGEODECOMP_CONFIG_H
#define LIBGEODECOMP_DEBUG_LEVEL 0
#define LIBGEODECOMP_WITH_CPP14 true
#ifdef __CUDACC__
/* nvcc translation units (tests/facade/generic_test.cu): enables the reference's __host__ __device__ paths */
#define LIBGEODECOMP_WITH_CUDA true
#endif
#ifdef _OPENMP
#define LIBGEODECOMP_WITH_THREADS true
#endif
#endif
