#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "tilted" 2>&1 | tail -30 | tee gpurun_out/r2C_pytest.log
