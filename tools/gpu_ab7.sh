#!/bin/bash
set -u
echo "== nostream"; FW_LIB_PATH=$PWD/firework_b200/libfw_nostream.so python tools/quick_bench.py 2>&1 | tail -9
echo "== stream"; python tools/quick_bench.py 2>&1 | tail -9
