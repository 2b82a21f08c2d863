#!/bin/bash
# parity + primitive tests of a variant build (copied over the main library for the tests), then same-box A/B
v=${1:-T3}
mkdir -p gpurun_out
cp computing-conjugate-points_b200/libccp_b200.so /tmp/main.so
cp computing-conjugate-points_b200/libccp_b200_$v.so computing-conjugate-points_b200/libccp_b200.so
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_oracle_primitives.py tests/test_subdivision.py -m gpu -q 2>&1 | tail -2
cp /tmp/main.so computing-conjugate-points_b200/libccp_b200.so
bash scripts/gpu_ab2.sh s4m_$v $v
