#!/bin/bash
for rep in 1 2; do
for v in "1 4 2" "0 4 2" "1 2 2" "1 4 1"; do set -- $v
echo "== cfg1 GRAPH=$1 STAGING=$2 EPI_SETS=$3"; VINCE_B200_GRAPH=$1 VINCE_B200_STAGING=$2 VINCE_B200_EPI_SETS=$3 timeout 600 python bench.py --config 1 --steps 30 --warmup 5 --profile-only 2>&1 | tail -1
done
done
