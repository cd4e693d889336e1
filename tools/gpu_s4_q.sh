#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_train_step.py tests/test_gpu_eval_sampler_batch.py -m gpu -q -k "c5_full or zero_steps or c4_item" 2>&1 | tail -25
timeout 300 python bench.py --steps 2000 --warmup 100 --no-eval --cpu-steps 20 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read()); print(j['clocks'], j['value'])"
