#!/bin/bash
# Tuning builds of liblbfgsb200.so with other tile shapes: build/variants/lib_<name>.so (select with LBFGSB200_SO).
set -e
cd "$(dirname "$0")/../rust_lbfgs_b200/csrc"
build() {  # name U MINBLOCKS UH [THREADS]
  mkdir -p ../../build/variants
  make -s -j8 OUT=../../build/variants/lib_$1.so OBJDIR=../../build/obj_$1 EXTRA="-DLB_U=$2 -DLB_MINBLOCKS=$3 -DLB_UH=$4 -DLB_THREADS=${5:-256} -DLB_UT=${6:-$2} ${EXTRA2:-}" 2>&1 | grep -i "error" || true
}
for v in "$@"; do
  case $v in
    u8b2) build u8b2 8 2 4;;
    u4b2) build u4b2 4 2 2;;
    u6b2) build u6b2 6 2 3;;
    u2b4) build u2b4 2 4 1;;
    u4b3) build u4b3 4 3 2;;
    u8b1) build u8b1 8 1 4;;
    u12b1) build u12b1 12 1 6;;
    t512u4b1) build t512u4b1 4 1 2 512;;
    t512u2b2) build t512u2b2 2 2 1 512;;
    t128u8b4) build t128u8b4 8 4 4 128;;
    t128u16b2) build t128u16b2 16 2 8 128;;
    ld1st0) EXTRA2="-DLB_LD_POLICY=1 -DLB_ST_POLICY=0" build ld1st0 8 2 4 256 6;;
    ld0st1) EXTRA2="-DLB_LD_POLICY=0 -DLB_ST_POLICY=1" build ld0st1 8 2 4 256 6;;
    ld1st1) EXTRA2="-DLB_LD_POLICY=1 -DLB_ST_POLICY=1" build ld1st1 8 2 4 256 6;;
    ld0st2) EXTRA2="-DLB_LD_POLICY=0 -DLB_ST_POLICY=2" build ld0st2 8 2 4 256 6;;
    ld2st0) EXTRA2="-DLB_LD_POLICY=2 -DLB_ST_POLICY=0" build ld2st0 8 2 4 256 6;;
    vA) build vA 6 1 3 256 5;;
    vB) build vB 7 1 4 256 6;;
    vC) build vC 8 1 5 256 7;;
    vD) build vD 10 1 4 256 10;;
    vE) build vE 8 1 4 256 12;;
    vF) build vF 9 1 4 256 9;;
    vG) build vG 8 1 4 256 3;;
    vH) build vH 8 1 4 256 4;;
    x-*) IFS=- read -r _ u uh ut <<< "$v"; build "$v" "$u" 1 "$uh" 256 "$ut";;   # x-<U>-<UH>-<UT>
    s-*) IFS=- read -r _ ld st <<< "$v"; EXTRA2="-DLB_LD_POLICY=$ld -DLB_ST_POLICY=$st" build "$v" 8 1 4 256 6;;   # s-<LD>-<ST>: cache policies, shipped tiles
    cg-*) IFS=- read -r _ u b <<< "$v"; EXTRA2="-DLB_CG_U=$u -DLB_CG_BLOCKS=$b" build "$v" 8 2 4 256 6;;   # cg-<U>-<CTAs per SM>: commit fused with pass A
    pm-*) IFS=- read -r _ b <<< "$v"; EXTRA2="-DLB_PM_BLOCKS=$b" build "$v" 8 2 4 256 6;;   # pm-<CTAs per SM>: the multi-step probe
    gram-*) IFS=- read -r _ u <<< "$v"; EXTRA2="-DLB_GRAM_U=$u" build "$v" 8 2 4 256 6;;   # gram-<U>: pass A of the compact direction, 3..5 older pairs
    p-*) IFS=- read -r _ u uh ut <<< "$v"; EXTRA2="-DLB_PROBE_PREFETCH=1" build "$v" "$u" 1 "$uh" 256 "$ut";;   # the same with the prefetching probe
  esac
done
ls -la ../../build/variants
