#!/bin/bash
# dev loop: run the GPU R1CS parity tests
timeout 900 python -m pytest tests/test_gpu_r1cs.py -x -q -m gpu 2>&1 | tail -30
