#!/bin/bash
MMF_TC_VARIANT=41 timeout -k 10 200 python -m pytest tests -m gpu -q -k "predict_measure and bf16" 2>&1 | tail -3
MMF_TC_VARIANT=41 timeout 120 python tools/time_step.py bf16x3 bf16 | cut -c1-75
