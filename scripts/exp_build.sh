#!/bin/bash
# Experiment helper: builds a variant of the library with extra nvcc flags.
# usage: scripts/exp_build.sh <name> [extra nvcc flags...]   → build/exp/lib_<name>.so
set -e
cd "$(dirname "$0")/.."
NAME=$1; shift
mkdir -p build/exp
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo --fmad=false -Xcompiler -fPIC -Xcompiler -ffp-contract=off -shared \
  -cudart static "$@" -o build/exp/lib_${NAME}.so dungeon_maps_b200/csrc/dm_api.cu dungeon_maps_b200/csrc/dm_project.cu dungeon_maps_b200/csrc/dm_labels.cu \
  dungeon_maps_b200/csrc/dm_flow.cu dungeon_maps_b200/csrc/dm_fuse.cu dungeon_maps_b200/csrc/dm_points.cu dungeon_maps_b200/csrc/dm_builder.cu dungeon_maps_b200/csrc/dm_params.cu dungeon_maps_b200/csrc/dm_ordered.cu
echo build/exp/lib_${NAME}.so
