#!/bin/bash
timeout -k 10 400 python -m pytest tests -m gpu -q -k "resample or normalize" 2>&1 | tail -4
timeout 120 python tools/time_step.py bf16x3 | cut -c1-160
