#!/bin/bash
# one short call: the face-BC / free-slip parity tests + a few existing tests that touch the changed export path
mkdir -p gpurun_out
export TXG_ASSUME_GPU=1
( time timeout 140 python -m pytest tests/test_zgpu_face_bcs.py "tests/test_gpu_parity.py::test_bubble_2d_golden" \
    "tests/test_gpu_parity.py::test_diagnostics_vs_oracle" "tests/test_gpu_parity.py::test_porous_srt_nonperiodic_box" \
    "tests/test_gpu_parity.py::test_porous_split_path_order4" \
    -q -m gpu --tb=short -p no:cacheprovider ) > gpurun_out/r1p_bc_tests.log 2>&1
tail -60 gpurun_out/r1p_bc_tests.log
