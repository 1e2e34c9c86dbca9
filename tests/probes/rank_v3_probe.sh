#!/bin/bash
# v2 vs v3 counting kernel: parity tests, then the bench's own per-kernel timers per variant.
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_rank.py tests/test_gpu_dropin.py -m gpu -x -q > gpurun_out/r2_v3_tests.log 2>&1
tail -5 gpurun_out/r2_v3_tests.log
for cfg in "DALI_RANK_V3=1" "DALI_RANK_V3_TIGHT=1" "DALI_RANK_V3_TIGHT=1 DALI_RANK_V3_THREADS=128"; do
  tag=$(echo "$cfg" | tr ' =' '__')
  env $cfg python bench.py --steps 20 --warmup 3 --no-c5 --no-modes --no-cpu-baseline > gpurun_out/r2_v3_$tag.json 2> gpurun_out/r2_v3_$tag.err
  python - "$tag" <<'PY'
import json,sys
tag=sys.argv[1]
for l in open(f'gpurun_out/r2_v3_{tag}.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print(tag, 'step', round(d['ms_per_step'],4), d['kernel_ms_per_step'], 'rank frac', d['roofline_rank_stage']['frac'],
              'c1', d.get('c1_market_vit',{}).get('ms_per_step'), d.get('c1_market_vit',{}).get('kernel_ms_per_step'),
              'c4', d.get('c4_fusion3',{}).get('ms_per_step'), 'mAP', d['mAP'])
PY
done
