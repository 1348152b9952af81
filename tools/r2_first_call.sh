#!/bin/bash
# First GPU call of round 2: `tools/r2_build_variants.sh && gpurun --timeout 1500 -- bash tools/r2_first_call.sh`.
# GPU tests of the product build, then the A/B of every staged variant (hashes must be identical), then the bench line.
set -u
O=gpurun_out
mkdir -p $O
( time python -m pytest tests -m gpu -x -q ) > $O/tests.log 2>&1
tail -3 $O/tests.log
bash tools/r2_render_ab.sh
bash tools/r2_precompute_ab.sh
python bench.py > $O/bench_default.json 2> $O/bench_default.err
cat $O/bench_default.json
