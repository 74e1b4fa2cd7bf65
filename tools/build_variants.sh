#!/bin/bash
# Builds tuning variants of libsph_b200 (same sources, different -D switches) next to the product library:
#   tools/build_variants.sh name1 "-DFLAG1 -DFLAG2" name2 "-D..." ...
# Select one at run time with SPH_B200_LIB=astrophysical-sph_b200/libsph_b200_<name>.so (tools/walk_tune.py takes it as
# an ENV=VALUE setting).  Not part of the product path.
set -e
cd "$(dirname "$0")/../astrophysical-sph_b200/csrc"
while [ $# -ge 2 ]; do
    make -j8 VARIANT="$1" EXTRA="$2" > /dev/null
    echo "built ../libsph_b200_$1.so with $2"
    shift 2
done
