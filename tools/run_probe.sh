#!/bin/bash
# runs every umma_probe test in its own process (a faulting variant must not take the others down)
mkdir -p gpurun_out
out=gpurun_out/umma_probe.log
: > $out
for t in 0 1 2 3 4 5 6 7 8 9 10 11 12 13 14 15 16 20 21 22 23 24 25; do
  timeout 30 tools/umma_probe $t >> $out 2>&1
  echo "exit code test $t: $?" >> $out
done
grep -E "RESULT|cycles/MMA|exit code|lanes holding" $out
