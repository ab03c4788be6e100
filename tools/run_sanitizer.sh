#!/bin/bash
# compute-sanitizer on the smallest configuration: one tool per gpurun call, each under its own timeout.
#   gpurun --timeout 900 -- 'tools/run_sanitizer.sh memcheck'      (then racecheck, synccheck, initcheck)
# Output: gpurun_out/sanitizer_<tool>.log; copy the summary lines into profiles/rNN_sanitizer.txt.
# The run is smoke() (8192 particles on a 400 x 400 map: rasterise, EDT, both score paths, normalise, resample) followed
# by the C++ host demo — every kernel family once, sizes the sanitizer finishes in minutes.
set -u
tool=${1:-memcheck}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 800 compute-sanitizer --tool "$tool" --error-exitcode 9 --log-file "gpurun_out/sanitizer_${tool}.log" \
  python -c "import __graft_entry__ as g; g.smoke()" > "gpurun_out/sanitizer_${tool}.stdout" 2>&1
echo "exit $? ($tool)"; tail -5 "gpurun_out/sanitizer_${tool}.log"
