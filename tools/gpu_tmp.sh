#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "permute or golden or transpose or c_abi" 2>&1 | tail -5
