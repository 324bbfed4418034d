#!/bin/bash
# Build a variant of libsvdpp.so for in-process A/B (tools/ab_lib.py):
#   tools/build_variant.sh <name> [<file.cu> ...]   -> ab_libs/libsvdpp_<name>.so
# Each given .cu replaces the same-named file of csrc/ in a scratch copy; the tree itself is untouched.
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
NAME="$1"; shift
T="$(mktemp -d)"
mkdir -p "$T/pkg" "$T/include" "$ROOT/ab_libs"
cp "$ROOT"/include/*.h "$T/include/"
mkdir -p "$T/pkg/csrc"
cp "$ROOT"/video-diffusion-pipeline-parallel_b200/csrc/{*.cu,*.cuh,*.h,Makefile} "$T/pkg/csrc/"
for f in "$@"; do cp "$f" "$T/pkg/csrc/$(basename "${f%%@*}")"; done
make -C "$T/pkg/csrc" -j6 >/dev/null
cp "$T/pkg/csrc/libsvdpp.so" "$ROOT/ab_libs/libsvdpp_$NAME.so"
rm -rf "$T"
echo "built ab_libs/libsvdpp_$NAME.so"
