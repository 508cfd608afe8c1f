#!/bin/bash
# Build a variant of libercgraph.so for A/B kernel experiments: tools/build_variant.sh <name> [git-rev|WORK] [EXTRA nvcc flags...]
# -> variants/<name>.so (git-ignored, travels to the GPU box); use with ERCG_LIB_PATH=variants/<name>.so
set -e
name=$1; rev=${2:-WORK}; shift; shift || true
root=$(cd "$(dirname "$0")/.." && pwd)
pkg=emotion-recognition-in-conversation_b200
tmp=/tmp/ercg_variant_$name
rm -rf $tmp; mkdir -p $tmp/$pkg $tmp/include
if [ "$rev" = WORK ]; then
  cp -r $root/$pkg/csrc $tmp/$pkg/; cp $root/include/*.h $tmp/include/
else
  (cd $root && git archive $rev $pkg/csrc include) | tar -x -C $tmp
fi
rm -rf $tmp/$pkg/csrc/build
make -C $tmp/$pkg/csrc -j 16 EXTRA="$*" > $tmp/build.log 2>&1 || { tail -30 $tmp/build.log; exit 1; }
mkdir -p $root/variants; cp $tmp/$pkg/libercgraph.so $root/variants/$name.so
echo "built variants/$name.so ($rev $*)"
