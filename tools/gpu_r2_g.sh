#!/usr/bin/env bash
set -u
out=gpurun_out/r2g
mkdir -p "$out"
t() { name=$1; shift; echo "== $name" | tee -a "$out/steps.log"; timeout 200 "$@" > "$out/$name.log" 2>&1; echo "rc=$? $(grep -h 'bit-exact\|illegal\|passed\|failed' "$out/$name.log" | head -1 | cut -c1-120)" | tee -a "$out/steps.log"; }
t py_env0 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "banded_pooled and env0"
MAREX_POOL_TMA=0 t py_env0_notma python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "banded_pooled and env0"
t r_noyear python tools/repro_ring2.py 0 13 70 2001-01-01
t r_year python tools/repro_ring2.py 1 13 70 2001-01-01
t r_year_64 python tools/repro_ring2.py 1 16 64 2001-01-01
t r_year_short python tools/repro_ring2.py 1 13 70 1996-01-01
MAREX_POOL_MARGIN=200 t r_year_margin python tools/repro_ring2.py 1 13 70 2001-01-01
