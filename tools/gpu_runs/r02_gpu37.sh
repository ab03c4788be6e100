#!/bin/bash
set -u
O=gpurun_out
python -m pytest tests -q -m gpu -x > $O/r02_pytest_gpu.txt 2>&1; echo "pytest rc $?"; tail -2 $O/r02_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_smoke.txt 2>&1; echo "smoke rc $?"; tail -1 $O/r02_smoke.txt
