#!/bin/bash
mkdir -p gpurun_out
PF_BLOCKING_SYNC=1 timeout 55 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-200
timeout 50 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-200
