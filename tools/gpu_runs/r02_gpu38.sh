#!/bin/bash
set -u
python -m pytest tests/test_gpu_bench_configs.py -x -q -m gpu -k "seven_classes" 2>&1 | tail -4
