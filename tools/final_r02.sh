#!/bin/bash
# final verification of the round-2 tree: every GPU test, smoke(), and the bench lines of the single-GPU workloads
O=gpurun_out
timeout 2000 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > $O/r02z_bench_default.json 2>/dev/null
python bench.py --workload flame > $O/r02z_bench_flame.json 2>/dev/null
python bench.py --workload latency1 --no-cpu-baseline > $O/r02z_bench_latency1.json 2>/dev/null
ls -la $O/r02z_*
