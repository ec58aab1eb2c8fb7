#!/bin/bash
# round 2, run AI: full GPU suite + smoke on the final tree
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q --timeout 100 -p no:cacheprovider > gpurun_out/r2ai_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2ai_pytest.log
grep -E "passed|failed|FAILED|rc=|^E  " gpurun_out/r2ai_pytest.log | tail -12
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
