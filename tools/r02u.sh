#!/bin/bash
# round-2 call U: RKS branch tests + full GPU suite
cd "$(dirname "$0")/.."
o=gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > $o/r02u_pytest.log 2>&1; tail -5 $o/r02u_pytest.log
