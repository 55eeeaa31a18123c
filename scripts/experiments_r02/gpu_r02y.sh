#!/bin/bash
for d in 0 1 8; do echo "== NCE_DEBUG=$d"; VINCE_B200_NCE_DEBUG=$d python tests/nce_kernel_probe.py 2>&1 | head -3; done
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
