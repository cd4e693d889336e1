#!/bin/bash
timeout 1500 python -m pytest tests/test_gpu_sampling_trainers.py -m gpu -q 2>&1 | tail -40
