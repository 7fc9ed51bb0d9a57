#!/bin/bash
for lib in libnw_sm100.so libnw_sm100_dbg4.so libnw_sm100_dbg5.so libnw_sm100.so; do echo "== $lib"; NW_B200_LIB=$PWD/nwhead_b200/$lib python tools/probe_perf.py 4096,1280000,256,1000 4096,1280000,512,1000 | awk '{print $1,$2,$3,$4,$5,$6,$7,$8,$9,$10,$11,$12,$13,$14,$15,$16,$17,$18,$19}'; done
