#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "thin or share_one or dataset_level or golden or streamed or partial_cover" 2>&1 | tail -15
