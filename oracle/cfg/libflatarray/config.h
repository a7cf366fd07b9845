/* Hand-written stand-in for lib/libflatarray/CMakeLists.txt:211's generated header. */
#ifndef LIBFLATARRAY_CONFIG_H
#define LIBFLATARRAY_CONFIG_H
#define LIBFLATARRAY_WITH_CPP14 true
#ifdef __CUDACC__
#define LIBFLATARRAY_WITH_CUDA true
#endif
#endif
