#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/fe_tests.log 2>&1; echo "tests rc=$?"; tail -30 gpurun_out/fe_tests.log
