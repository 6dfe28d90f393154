#!/bin/bash
# A/B builds of the closure kernel on the GPU box: bash tools/rows_ab.sh "-DTRL_ROWS_SPECIAL=0" "-DTRL_ROWS_SPECIAL=1" ...
for cfg in "$@"; do
  TRL_NVCC_EXTRA="$cfg" python -m tetris_reinforcement_learning_b200.build --force > /dev/null 2>&1
  echo "== $cfg"
  python tools/movegen_bench.py 100000 2>&1 | grep -E "n_calls=  700000 kernel=rows   out=mask|thread vs rows" | tail -2
done
python -m tetris_reinforcement_learning_b200.build --force > /dev/null 2>&1
