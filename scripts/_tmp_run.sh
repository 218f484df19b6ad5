#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_pme.py -q -m gpu -x 2>&1 | grep -E "passed|failed|^E  |assert" | head -20
