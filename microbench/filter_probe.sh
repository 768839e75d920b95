#!/bin/bash
# sweep of microbench/filter_probe.cu (built into microbench/bin/ by `make microbench`)
B=microbench/bin/filter_probe
$B 0 0 -1
for t in 1070 615 512; do $B $t 77 0 35; done
for f in 38 77 100; do $B 615 $f 1 35; $B 615 $f 1 35 1; done
for m in 2 3 4; do for p in 0 1; do $B 615 77 $m 35 $p; done; done
for m in 2 3 4; do $B 615 77 $m 10 1; done
$B 1070 64 3 35 1
$B 615 38 3 35 1
$B 615 38 3 50 1
