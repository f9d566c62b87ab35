#!/bin/bash
# SASS listings of the stage kernels (K = 11 instantiations) from the built library -> profiles/sass/
set -eu
cd "$(dirname "$0")/.."
LIB=bayes_od_rc_b200/lib/libbayesod.so
OUT=profiles/sass
mkdir -p $OUT
rm -f $OUT/*.sass
dump() {  # $1 = regex on the mangled name, $2 = output name
  fn=$(cuobjdump -elf $LIB 2>/dev/null | grep -o "\.text\.[A-Za-z0-9_]*" | sed 's/^\.text\.//' | grep -E "$1" | head -1)
  [ -z "$fn" ] && { echo "no kernel matches $1"; return; }
  cuobjdump -sass -fun "$fn" $LIB 2>/dev/null > $OUT/$2.sass
  echo "$2: $(grep -c ';' $OUT/$2.sass) lines; UBLKCP=$(grep -c UBLKCP $OUT/$2.sass || true) SYNCS=$(grep -c SYNCS $OUT/$2.sass || true) REDUX=$(grep -c REDUX $OUT/$2.sass || true) FFMA=$(grep -c FFMA $OUT/$2.sass || true) FFMA2=$(grep -c FFMA2 $OUT/$2.sass || true)"
}
dump 'k1_moments_pipe_kernelILi11ELi5' k1_moments_pipe_kernel_K11    # the unrolled five-stage round of the bench workload (FFMA2 / FADD2)
dump 'k1_moments_pipe_kernelILi11ELi0' k1_moments_pipe_kernel_K11_generic
dump 'scan_tiles_kernel' scan_tiles_kernel
dump 'k2_posterior_kernelILi11' k2_posterior_kernel_K11
dump 'k3_softnms_kernel' k3_softnms_kernel
dump 'k4_fusion_kernelILi11' k4_fusion_kernel_K11
dump 'prefilter_select_kernel' prefilter_select_kernel
dump 'val_filter_kernelILi11' val_filter_kernel_K11
dump 'pdq_table_kernel' pdq_table_kernel
dump 'pdq_sum_kernel' pdq_sum_kernel
dump 'pdq_heatmap_kernelE' pdq_heatmap_kernel
dump 'pdq_roi_kernel' pdq_roi_kernel
dump 'k1_moments_pipe_kernelILi8ELi5' k1_moments_pipe_kernel_K8
dump 'k1_moments_pipe_kernelILi4ELi5' k1_moments_pipe_kernel_K4
dump 'entropy_kernel' entropy_kernel
dump 'mue_match_kernel' mue_match_kernel
