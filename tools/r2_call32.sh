#!/bin/bash
timeout -k 5 100 python -m pytest tests/test_e2e_gpu.py -m gpu -x -q -k "two_phase" 2>&1 | tail -5 | cut -c1-300
