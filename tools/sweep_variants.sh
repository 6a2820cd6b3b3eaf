#!/bin/bash
# timing (and optionally parity) of the tensor-core chain kernel variants
for v in ${VARIANTS:-41 51}; do
  for lock in ${LOCKS:-0}; do
    echo "== MMF_TC_VARIANT=$v MMF_TC_LOCK=$lock"
    if [ -n "$PARITY" ]; then
      MMF_TC_VARIANT=$v MMF_TC_LOCK=$lock timeout -k 10 400 python -m pytest tests -m gpu -q -x -k "predict or step or bptt or heads" 2>&1 | tail -3 | cut -c1-300
    fi
    MMF_TC_VARIANT=$v MMF_TC_LOCK=$lock timeout 120 python tools/time_step.py bf16x3 bf16 | cut -c1-120
  done
done
