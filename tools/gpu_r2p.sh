#!/bin/bash
# Round-2 session P: compute-sanitizer memcheck (all entry points, s..f shells) and racecheck (J/K kernels) on small molecules,
# the new primitive-cutoff margin test, smoke() with its Hessian check.
TAG=${TAG:-r2p}
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_case.py hf_tz h2o > gpurun_out/${TAG}_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/${TAG}_memcheck.log
SAN_MODE=jk timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/sanitize_case.py hf_tz > gpurun_out/${TAG}_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/${TAG}_racecheck.log
timeout 600 python -m pytest tests -m gpu -q -x -k "primitive_cutoff or hess or golden" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/${TAG}_smoke.log
