#!/bin/bash
# A/B of the staged sky-evaluation variants against the product build, one GPU pass (run under gpurun AFTER building
# the variants here with tools/r2_build_variants.sh; the .so files travel).
# For each variant: tools/render_ab.py writes one SHA-256 per output of a 41-view sweep and the device time of the 4K
# frames; a variant is adopted only if its hash file is identical to the product build's and it is faster.
set -u
O=gpurun_out
mkdir -p $O
python tools/render_ab.py $O/render_ab_product.txt > $O/render_ab_product.log 2>&1
for v in build/variants/render_*.so; do
    n=$(basename $v .so)
    FUZZYBLUE_B200_LIB=$PWD/$v python tools/render_ab.py $O/render_ab_$n.txt > $O/render_ab_$n.log 2>&1
    if cmp -s $O/render_ab_product.txt $O/render_ab_$n.txt; then same=identical; else same=DIFFERENT; fi
    echo "$n: hashes $same; product $(tail -1 $O/render_ab_product.log); variant $(tail -1 $O/render_ab_$n.log)"
done
