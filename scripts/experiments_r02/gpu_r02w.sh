#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
   -k regex:'infonce_main|ema_enqueue' -c 4 -o gpurun_out/r02_top_nce -f \
   python bench.py --config 2 --steps 2 --warmup 3 --profile-only > gpurun_out/ncu_top.log 2>&1
tail -1 gpurun_out/ncu_top.log
ncu -i gpurun_out/r02_top_nce.ncu-rep --page raw --csv > /tmp/top_nce.csv 2>/dev/null; python scripts/ncu_reduce.py /tmp/top_nce.csv gpurun_out/r02_top_nce_set_full.csv
rm -f gpurun_out/r02_top_nce.ncu-rep
cut -d, -f2,5,6,7 gpurun_out/r02_top_nce_set_full.csv | head -8
python tests/nce_host_probe.py ResNet50 2>&1 | head -1
python tests/nce_host_probe.py ResNet18 2>&1 | head -1
