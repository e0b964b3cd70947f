#!/usr/bin/env bash
set -u
out=gpurun_out/r2i
mkdir -p "$out"
t() { name=$1; shift; echo "== $name" | tee -a "$out/steps.log"; timeout 300 "$@" > "$out/$name.log" 2>&1; echo "rc=$? $(grep -h 'bit-exact\|illegal\|passed\|failed' "$out/$name.log" | head -1 | cut -c1-120)" | tee -a "$out/steps.log"; }
t base python tools/repro_ring2.py 1 13 70 2001-01-01
t plain python tools/repro_ring.py 16 64 11 11 1
