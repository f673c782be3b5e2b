#!/bin/bash
# same-box A/B of two builds of the library:  tools/ab.sh "C2:50 C3:200" [lib ...]   (default: the in-tree build vs lib/libsb2gpu_base.so)
cd "$(dirname "$0")/.."
specs=${1:-C2:50}; shift
libs=${@:-"starbeast2_b200/lib/libsb2gpu_base.so starbeast2_b200/lib/libsb2gpu.so"}
for rep in 1 2; do
for lib in $libs; do
  echo "== $lib (rep $rep)"
  SB2_LIB=$PWD/$lib tools/gpu_check.sh --notest "$specs"
done
done
