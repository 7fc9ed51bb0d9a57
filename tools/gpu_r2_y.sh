#!/bin/bash
# round-2 GPU call Y (1 GPU): forward + aux tests (new: metadata paths, influence shapes)
timeout 900 python -m pytest tests/test_gpu_forward.py tests/test_gpu_aux.py -q -m gpu 2>&1 | tail -25
