#!/bin/bash
# whole GPU test suite, verbose about failures
set -u
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -s "$@" 2>&1 | grep -v "^$" | tail -60 | tee gpurun_out/tests_last.log
