#!/bin/sh
# TEST INFRASTRUCTURE ONLY.  Prints the first C++ compiler among the arguments that can build AND link an OpenMP program.
# (This image exports CXX=/opt/gcc/bin/g++, a wrapper that has no libgomp.spec; the environment's CXX is not trusted.)
t=$(mktemp -d)
trap 'rm -rf "$t"' EXIT
printf '#include <omp.h>\nint main(){return omp_get_max_threads()>0?0:1;}\n' > "$t/p.cpp"
for c in "$@"; do
  command -v "$c" >/dev/null 2>&1 || continue
  if "$c" -fopenmp "$t/p.cpp" -o "$t/p" >/dev/null 2>&1; then echo "$c"; exit 0; fi
done
exit 0
