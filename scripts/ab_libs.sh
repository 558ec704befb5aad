#!/bin/bash
# A/B a set of prebuilt library variants (build/libs/lib_*.so) on one command; restores the default library.
cp mmtg_b200/lib/libmmtg_b200.so /tmp/lib_default.so
for f in build/libs/lib_*.so; do
  cp $f mmtg_b200/lib/libmmtg_b200.so
  echo "== $f"; "$@" 2>&1 | tail -2
done
cp /tmp/lib_default.so mmtg_b200/lib/libmmtg_b200.so
