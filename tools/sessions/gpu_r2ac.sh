timeout 900 python -m pytest tests/test_gpu_fused.py -m gpu -x -q 2>&1 | grep -E "assert|Error|error|fused.py" | head -30
