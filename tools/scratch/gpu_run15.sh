timeout -k 10 600 python -m pytest tests/test_training_glue.py -m gpu -q -x -k "hoists" 2>&1 | grep -v "^  File\|Warning" | tail -40
