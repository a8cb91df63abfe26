#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "verif or full_proof or fiat or cache or graph or cpp_header or rejected or batch or generate_then" 2>&1 | tail -2
