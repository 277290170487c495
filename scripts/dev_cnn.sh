#!/bin/bash
# CNN A/B: GEMM K-block variants + tests
python -m pytest tests/test_cnn_gpu.py -x -q 2>&1 | tail -4
echo "BK=64"; EPOS_GEMM_BK=64 python scripts/dev_time.py 8 21 5
echo "BK=32"; EPOS_GEMM_BK=32 python scripts/dev_time.py 8 21 5
echo "default"; python scripts/dev_time.py 8 21 5
