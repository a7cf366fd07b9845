#!/bin/bash
# Device code of two revisions side by side, no GPU needed: compiles every translation unit of libb200geo.so and the
# generic-path test (tests/facade/generic_test.cu) of revision $1 (default: the last GPU-verified commit of round 1) and of
# the working tree to sm_100a cubins and diffs their SASS (addresses, encodings and the anonymous-namespace hash stripped).
# 0 differing lines = the kernels that were measured and verified on the B200 are the kernels in this tree.
#   tools/sass_diff.sh [revision]
REV=${1:-5929b35}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
REF=${REF:-/root/reference}
WORK=$(mktemp -d)
mkdir -p $WORK/old && git -C $ROOT archive $REV include tests/facade oracle/models oracle/cfg libgeodecomp_b200/csrc | tar -x -C $WORK/old
strip() { cuobjdump -sass $1 2>/dev/null | sed 's#/\*[0-9a-f]\{4\}\*/##; s#/\* 0x[0-9a-f]* \*/##; s/_GLOBAL__N__[0-9a-f]\{8\}/_GLOBAL__N__X/g'; }
for f in gol gol_bits grid group jacobi jacobi_tb lbm nbody region; do
  (
    flags="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -w -cubin"
    case $f in lbm|nbody) flags="$flags -fmad=false";; esac
    nvcc $flags $WORK/old/libgeodecomp_b200/csrc/$f.cu -o $WORK/old_$f.cubin 2>/dev/null
    nvcc $flags $ROOT/libgeodecomp_b200/csrc/$f.cu -o $WORK/new_$f.cubin 2>/dev/null
    strip $WORK/old_$f.cubin > $WORK/old_$f.sass; strip $WORK/new_$f.cubin > $WORK/new_$f.sass
    echo "csrc/$f.cu: $(grep -c Function $WORK/new_$f.sass) kernels, $(wc -l < $WORK/new_$f.sass) SASS lines, $(diff $WORK/old_$f.sass $WORK/new_$f.sass | wc -l) differing"
  ) &
done
if [ -d $REF/src/libgeodecomp ]; then
  (
    flags="-std=c++14 -O2 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -w -cubin -I$REF/src -I$REF/lib/libflatarray/include"
    (cd $WORK/old/tests/facade && nvcc $flags -I../../oracle/cfg -I../../oracle -I../../include generic_test.cu -o $WORK/old_generic.cubin 2>/dev/null)
    (cd $ROOT/tests/facade && nvcc $flags -I../../oracle/cfg -I../../oracle -I../../include generic_test.cu -o $WORK/new_generic.cubin 2>/dev/null)
    strip $WORK/old_generic.cubin > $WORK/old_generic.sass; strip $WORK/new_generic.cubin > $WORK/new_generic.sass
    echo "tests/facade/generic_test.cu: $(grep -c Function $WORK/new_generic.sass) kernels, $(wc -l < $WORK/new_generic.sass) SASS lines, $(diff $WORK/old_generic.sass $WORK/new_generic.sass | wc -l) differing"
  ) &
fi
wait
rm -rf $WORK
