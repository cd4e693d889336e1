#!/bin/bash
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -E "^E  |Error|FAILED|passed|failed" | head -40
