#!/bin/bash
# Builds the library of another commit into tools/build/libtinyvc_b200_<name>.so for same-box A/B runs (TVC_LIB=...).
# usage: tools/build_prev_lib.sh <commit> <name>
set -e
C=${1:-HEAD}; N=${2:-prev}
D=$(mktemp -d)
git -C "$(dirname "$0")/.." archive "$C" tinyvc_b200 include | tar -x -C "$D"
( cd "$D" && python -m tinyvc_b200.build >/dev/null )
mkdir -p "$(dirname "$0")/build"
cp "$D/tinyvc_b200/libtinyvc_b200.so" "$(dirname "$0")/build/libtinyvc_b200_$N.so"
rm -rf "$D"
echo "built tools/build/libtinyvc_b200_$N.so from $C"
