#!/bin/bash
python tests/conv_bench.py --filter "stem" --iters 9
python tests/conv_bench.py --filter "r18." --iters 7
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
for c in 1 2; do echo "== bench cfg$c"; timeout 600 python bench.py --config $c --steps 30 --warmup 5 --profile-only 2>&1 | tail -1; done
