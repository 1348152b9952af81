#!/bin/bash
# A/B of the staged precompute variants against the product build, one GPU pass (run under gpurun AFTER building the
# variants here with tools/r2_build_variants.sh; the .so files travel).
# tools/precompute_ab.py writes one SHA-256 per table for three dims and the per-stage device times; a bit-identical
# variant is adopted only if its hash file equals the product build's and its stage is faster.
set -u
O=gpurun_out
mkdir -p $O
python tools/precompute_ab.py $O/precompute_ab_product.txt 20 > $O/precompute_ab_product.log 2>&1
echo "product: $(tail -1 $O/precompute_ab_product.log)"
for v in build/variants/pre_*.so; do
    n=$(basename $v .so)
    FUZZYBLUE_B200_LIB=$PWD/$v python tools/precompute_ab.py $O/precompute_ab_$n.txt 20 > $O/precompute_ab_$n.log 2>&1
    if cmp -s $O/precompute_ab_product.txt $O/precompute_ab_$n.txt; then same=identical; else same=DIFFERENT; fi
    echo "$n: hashes $same; $(tail -1 $O/precompute_ab_$n.log)"
done
