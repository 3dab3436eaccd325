set -x
python -m pytest tests -m gpu -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python bench_ops.py --workload all --iters 30 --json gpurun_out/r1_bench_ops.jsonl 2>&1 | grep -v "^\[" > gpurun_out/r1_bench_ops.txt
python bench_ops.py --workload ops2 --iters 30 --json gpurun_out/r1_bench_ops2.jsonl 2>&1 | grep -v "^\[" >> gpurun_out/r1_bench_ops.txt
python bench.py --steps 10 --warmup 3 > gpurun_out/r1_bench_c2_n1.json 2> gpurun_out/bench_err.txt; tail -c 300 gpurun_out/r1_bench_c2_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches_bench_c2.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launch.log 2>&1; tail -1 gpurun_out/ncu_launch.log | cut -c1-120
timeout 300 ncu --set full --clock-control none --import-source on -k regex:staged -s 1 -c 1 -o gpurun_out/r1_warp_staged_v4 -f python profiles/_c3_once.py > gpurun_out/ncu_ws.log 2>&1; tail -1 gpurun_out/ncu_ws.log
