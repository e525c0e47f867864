#!/bin/bash
# GPU-box final verification: -m gpu suite (incl. result-assembly test), smoke, both bench lines.
mkdir -p gpurun_out
( time timeout 400 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider ) > gpurun_out/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest.log
( time timeout 120 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1
( timeout 200 python bench.py --steps 20 --warmup 3 ) > gpurun_out/bench_knlm.json 2> gpurun_out/bench_knlm.err
( timeout 200 python bench.py --model cong --steps 20 --warmup 3 --cpu-sample 8192 ) > gpurun_out/bench_cong.json 2> gpurun_out/bench_cong.err
( timeout 200 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
tail -n 25 gpurun_out/pytest.log | cut -c1-400; tail -n 3 gpurun_out/smoke.log
cut -c1-200 gpurun_out/bench_knlm.json; cut -c1-200 gpurun_out/bench_cong.json; cut -c1-200 gpurun_out/bench_ref.json
