#!/bin/bash
# round-2 last GPU call: the whole GPU suite and smoke() on the final commit
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02z_smoke.log 2>&1; echo "smoke rc=$?" > gpurun_out/r02z_rc.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02z_tests.log 2>&1; echo "suite rc=$?" >> gpurun_out/r02z_rc.txt
cat gpurun_out/r02z_rc.txt; tail -2 gpurun_out/r02z_smoke.log | cut -c1-300; tail -3 gpurun_out/r02z_tests.log
