#!/bin/bash
# A/B builds for the end-to-end path on the GPU box: bash tools/e2e_ab.sh "-DTRL_HOST_CHUNK_LOG2=17" "-DTRL_HOST_CHUNK_LOG2=18" ...
for cfg in "$@"; do
  TRL_NVCC_EXTRA="$cfg" python -m tetris_reinforcement_learning_b200.build --force > /dev/null 2>&1
  python bench.py --no-selfplay --no-cpu-baseline --steps 5 > /tmp/e2e_ab.json 2>/dev/null
  python -c "
import json; d=json.load(open('/tmp/e2e_ab.json')); print('$cfg', 'device %.3e e2e %.3e masks %.3e' % (d['value'], d['e2e']['value'], d['e2e']['as_bit_packed_masks']['value']), d['e2e']['matches_device_run'])"
done
python -m tetris_reinforcement_learning_b200.build --force > /dev/null 2>&1
