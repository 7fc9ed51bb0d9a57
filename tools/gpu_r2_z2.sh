#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_forward.py tests/test_gpu_scale.py tests/test_gpu_nwnet.py tests/test_gpu_dropin.py -q -x -m gpu 2>&1 | tail -2
python tools/probe_cfg4_launches.py
