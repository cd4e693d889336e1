#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_end_to_end.py -x -q -k meanpool_graph 2>&1 | grep -E "^E|Error|assert" | head -30
