#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12 | cut -c1-400 | tee gpurun_out/r2K_pytest.log
