#!/bin/bash
# A/B builds of libespic_b200.so with development switches (-DESPIC_VAR_*), into variants/ (git-ignored,
# travels to the GPU box).  usage: scripts/build_variants.sh name "-DESPIC_VAR_CARRY=0 ..." [name flags ...]
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p "$ROOT/variants"
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  tmp=$(mktemp -d)
  mkdir -p "$tmp/pkg/csrc" "$tmp/include"
  cp "$ROOT"/espic_hole_tracking_b200/csrc/*.cu "$ROOT"/espic_hole_tracking_b200/csrc/*.cuh "$ROOT"/espic_hole_tracking_b200/csrc/Makefile "$tmp/pkg/csrc/"
  cp "$ROOT"/include/espic_b200.h "$tmp/include/"
  make -C "$tmp/pkg/csrc" -j8 NVCC="/usr/local/cuda/bin/nvcc $flags" > "$tmp/build.log" 2>&1 || { tail -20 "$tmp/build.log"; exit 1; }
  cp "$tmp/pkg/libespic_b200.so" "$ROOT/variants/libespic_$name.so"
  grep -A2 "k_moveILb1ELb1ELb1E" "$tmp/pkg/csrc/mover.ptxas.log" | grep "spill\|Used" | tr '\n' ' ' | sed "s/^/$name: /; s/ptxas info    ://; s/  */ /g"; echo
  rm -rf "$tmp"
done
