#!/bin/bash
# re-entry check: whole GPU suite, smoke, 1-GPU bench line of the current head
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -8 ) 2>&1 | tee gpurun_out/gputests_33.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 | tee gpurun_out/smoke_33.log
( time python bench.py > gpurun_out/bench_33.json 2> gpurun_out/bench_33.err ) 2>&1 | tail -3
tail -c 3000 gpurun_out/bench_33.json
