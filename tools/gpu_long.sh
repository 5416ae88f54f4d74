#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -q -x -k long_sequence --tb=short 2>&1 | tail -n 25
timeout 300 python tools/long_sequences.py gpurun_out/long_sequences.jsonl 2>&1 | tail -n 14
