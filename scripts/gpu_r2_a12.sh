#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "iterate or projected or fused or sgap or tma or label_prop" 2>&1 | grep -E "passed|failed|Error|assert|^E " | head -30 | cut -c1-300
python scripts/fused_probe.py products 2>&1 | cut -c1-200
python scripts/fused_probe.py arxiv 2>&1 | cut -c1-200
