#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python -m pytest tests -m gpu -x -q ) > gpurun_out/last2_pytest.log 2>&1
echo done
