#!/bin/bash
# one short call: the face-BC / free-slip parity tests (TXG_ASSUME_GPU skips the torch import of conftest.py)
mkdir -p gpurun_out
export TXG_ASSUME_GPU=1
( time timeout 100 python -m pytest tests/test_zgpu_face_bcs.py -q -m gpu --tb=short -p no:cacheprovider ) > gpurun_out/r1r_bc_tests.log 2>&1
tail -40 gpurun_out/r1r_bc_tests.log
