#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 | cut -c1-300 | tee gpurun_out/r2T_pytest.log
