#!/bin/bash
# state of the head: whole GPU suite, smoke, reference arm, default bench line, ncu launch list of the same command
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4 ) 2>&1 | tee gpurun_out/gputests_38.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke_38.log
( time python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_38_ref.json 2> gpurun_out/bench_38_ref.err ) 2>&1 | tail -3
( time python bench.py > gpurun_out/bench_38.json 2> gpurun_out/bench_38.err ) 2>&1 | tail -3
python tools/bench_summary.py < gpurun_out/bench_38.json | head -12
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file gpurun_out/r02_launches_38.csv python bench.py --no-other > gpurun_out/ncu_launch38.log 2>&1
tail -2 gpurun_out/ncu_launch38.log | cut -c1-300
