#!/bin/bash
# timing (and optionally parity) of the tensor-core chain kernel variants (MMF_TC_VARIANT is read once per process)
for v in ${VARIANTS:-71 41}; do
  echo "== MMF_TC_VARIANT=$v"
  if [ -n "$PARITY" ]; then
    MMF_TC_VARIANT=$v timeout -k 10 600 python -m pytest tests -m gpu -q -x -k "predict or step or bptt or heads or fixture" 2>&1 | tail -3 | cut -c1-300
  fi
  MMF_TC_VARIANT=$v timeout 120 python tools/time_step.py bf16x3 bf16 | cut -c1-160
done
