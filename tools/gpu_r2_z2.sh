#!/bin/bash
for lib in libnw_sm100.so libnw_sm100_prev.so libnw_sm100.so libnw_sm100_prev.so; do echo "== $lib"; NW_B200_LIB=$PWD/nwhead_b200/$lib python tools/probe_perf.py 4096,1280000,256,1000 4096,1280000,512,1000 4096,1280000,384,1000 | cut -c1-130; done
