#!/bin/bash
for v in 41 31 21 32 42 22; do
  for h in 0 20000; do
    echo "variant=$v hint=$h"; MMF_TC_VARIANT=$v MMF_TC_WAIT_HINT_NS=$h timeout 120 python tools/time_step.py bf16x3 bf16 | cut -c1-75
  done
done
