#!/bin/bash
# Builds every variant staged for round 2 into build/variants/ (run HERE, before the gpurun call: the .so files travel).
set -e
cd "$(dirname "$0")/.."
rm -rf build/variants
tools/build_render_variant.sh render_magicfloor "-DFB_RENDER_MAGIC_FLOOR=1"
tools/build_render_variant.sh render_skysplit "-DFB_RENDER_SKY_SPLIT=1"
tools/build_render_variant.sh render_skysplit_magic "-DFB_RENDER_SKY_SPLIT=1 -DFB_RENDER_MAGIC_FLOOR=1"
tools/build_variant.sh pre_ms_diet "-DFB_MS_DIET=1" > /dev/null 2>&1 && echo built build/variants/pre_ms_diet.so
tools/build_variant.sh pre_ms_tpt2 "-DFB_MS_TPT2=1" > /dev/null 2>&1 && echo built build/variants/pre_ms_tpt2.so
tools/build_variant.sh pre_ms_tpt2_diet "-DFB_MS_DIET=1 -DFB_MS_TPT2=1" > /dev/null 2>&1 && echo built build/variants/pre_ms_tpt2_diet.so
tools/build_variant.sh pre_ss_tpt2 "-DFB_SS_TPT2=1" > /dev/null 2>&1 && echo built build/variants/pre_ss_tpt2.so
tools/build_variant.sh pre_density_rows "-DFB_DENSITY_ROWS=1" > /dev/null 2>&1 && echo built build/variants/pre_density_rows.so
tools/build_variant.sh pre_all "-DFB_DENSITY_ROWS=1 -DFB_SS_TPT2=1 -DFB_MS_DIET=1 -DFB_MS_TPT2=1" > /dev/null 2>&1 && echo built build/variants/pre_all.so
