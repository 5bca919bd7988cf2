#!/bin/bash
# Round-end style verification: all GPU tests, smoke(), the default bench line (with the CPU reference arm).
out=gpurun_out/${1:-final}; mkdir -p $out
(time timeout 1500 python -m pytest tests -m gpu -q --maxfail=10) > $out/tests_all.log 2>&1; grep -E "passed|failed|^FAILED|^ERROR" $out/tests_all.log | head -30
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > $out/bench.json 2> $out/bench.err; cat $out/bench.json; tail -3 $out/bench.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > $out/bench_ref.json 2> $out/bench_ref.err; cat $out/bench_ref.json | cut -c1-600
