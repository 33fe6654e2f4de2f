#!/bin/bash
# last GPU call of round 2 (about two minutes of budget left): the new seeded / sharded Datalog tests first, then the Datalog tests that
# cover the refactored driver, the timing script, then the rest of the suite for as long as the box lives. Everything is teed into
# gpurun_out/ as it happens so that a cut call still leaves its evidence.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
L=gpurun_out/r3e.log
: > $L
T0=$SECONDS
timeout 80 python -m pytest tests/test_gpu_sharded_fixpoint.py -m gpu -v -p no:cacheprovider 2>&1 | grep --line-buffered -v "^$" | tee -a $L | tail -45
echo "== new tests: $((SECONDS-T0)) s" | tee -a $L
timeout 70 python -m pytest tests/test_gpu_datalog.py tests/test_gpu_fuzz.py tests/test_gpu_rsp.py tests/test_gpu_fullsize.py::test_cfg4_closure_full_size -m gpu -q -p no:cacheprovider 2>&1 | tail -8 | tee -a $L
echo "== datalog regression: $((SECONDS-T0)) s" | tee -a $L
timeout 60 python scripts/seed_bench.py 2>&1 | tail -3 | tee gpurun_out/seed_bench_r3e.json | tee -a $L
echo "== seed bench: $((SECONDS-T0)) s" | tee -a $L
timeout 120 python -m pytest tests -m gpu -q -x -p no:cacheprovider --deselect tests/test_gpu_sharded_fixpoint.py --deselect tests/test_gpu_datalog.py --deselect tests/test_gpu_fuzz.py --deselect tests/test_gpu_rsp.py 2>&1 | tail -6 | tee -a $L
echo "== rest of the suite: $((SECONDS-T0)) s" | tee -a $L
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee -a $L
echo "== smoke: $((SECONDS-T0)) s" | tee -a $L
