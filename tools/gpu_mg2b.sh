#!/bin/bash
( time timeout 1200 python -m pytest tests/test_multi_gpu.py -m gpu -q ) 2>&1 | grep -E "delta_norm|passed|failed|Error|assert" | cut -c1-400 | tail -20
