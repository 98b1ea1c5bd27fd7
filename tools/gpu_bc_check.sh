#!/bin/bash
# one short call: the face-BC / free-slip / EOS parity tests + a few existing tests that touch the changed paths
mkdir -p gpurun_out
export TXG_ASSUME_GPU=1
( time timeout 150 python -m pytest tests/test_zgpu_face_bcs.py tests/test_zgpu_eos.py "tests/test_gpu_parity.py::test_bubble_2d_golden" \
    "tests/test_gpu_parity.py::test_porous_shan_chen_eos" "tests/test_gpu_parity.py::test_porous_mrt_minerals_body" \
    -q -m gpu --tb=short -p no:cacheprovider ) > gpurun_out/r1q_bc_eos_tests.log 2>&1
tail -60 gpurun_out/r1q_bc_eos_tests.log
