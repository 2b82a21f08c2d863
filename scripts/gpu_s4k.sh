#!/bin/bash
# parity tests of variant builds (CCP_LIB is honoured by scripts only, so the variant is copied over the main library for the test), then A/B
mkdir -p gpurun_out
cp computing-conjugate-points_b200/libccp_b200.so /tmp/main.so
cp computing-conjugate-points_b200/libccp_b200_T1.so computing-conjugate-points_b200/libccp_b200.so
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -2
cp /tmp/main.so computing-conjugate-points_b200/libccp_b200.so
bash scripts/gpu_ab2.sh s4k T1 T1G4 T1G4H2
