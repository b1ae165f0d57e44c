#!/bin/bash
# A/B timing of two builds of libsurs.so on ONE box (boxes differ by several per cent): alternates
#   A = $1 (e.g. experiments/_libsurs_scalar.so)   B = the in-tree build
# usage: scripts/ab_grid.sh A.so [precision ...]
A=$1; shift
PRECS=${@:-fp16}
for p in $PRECS; do
  for i in 1 2; do
    echo -n "A $p: "; GRID_REPS=7 SURS_LIB=$A timeout 120 python scripts/grid_once.py 512 $p 2>&1 | tail -1
    echo -n "B $p: "; GRID_REPS=7 timeout 120 python scripts/grid_once.py 512 $p 2>&1 | tail -1
  done
done
