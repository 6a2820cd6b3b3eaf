timeout 900 python -m pytest tests -m gpu -q -x -k "resample or normalize" 2>&1 | grep -E "^FAILED|Error|assert|differ" | head -12
