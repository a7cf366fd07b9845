#!/bin/bash
# A/B in one call on one box: old temporal-blocked kernel (state rotation) vs the phase-alternating one
mkdir -p gpurun_out
run() {  # $1 = library, rest = tune.py args
  lib=$1; shift
  python - "$lib" "$@" <<'PY'
import runpy, sys
import libgeodecomp_b200.capi as c
c._LIB_PATH = sys.argv[1]
sys.argv = ["tools/tune.py"] + sys.argv[2:]
runpy.run_path("tools/tune.py", run_name="__main__")
PY
}
NEW=$PWD/libgeodecomp_b200/libb200geo.so
OLD=$PWD/gpurun_out_ab_libb200geo_old.so
for rep in 1 2 3; do
  echo "old:"; run $OLD jacobi27 jacobi.tb=2 2>&1 | tail -1
  echo "new:"; run $NEW jacobi27 jacobi.tb=2 2>&1 | tail -1
done
for rep in 1 2; do
  echo "old:"; run $OLD jacobi7 jacobi.tb=4 2>&1 | tail -1
  echo "new:"; run $NEW jacobi7 jacobi.tb=4 2>&1 | tail -1
done
