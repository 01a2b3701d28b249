#!/bin/bash
cd "$(dirname "$0")/../.."
for mb in 6 8 10 12; do
  echo "=== PRE3_EVP_MINB=$mb"
  PRE3_EVP_MINB=$mb timeout 300 python bench.py --steps 20 --warmup 5 --no-other 2>/dev/null | python tools/bench_summary.py | sed -n 1,2p
done
timeout 200 tools/evalbench.bin 20000 454656 2>&1 | head -3
