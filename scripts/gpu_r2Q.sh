#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "filtered_rods" 2>&1 | tail -30 | cut -c1-400 | tee gpurun_out/r2Q_pytest.log
