# round 2, GPU trip A: tests, smoke, bench line, per-operator table, PCIe ceiling, launch list, ncu captures of the evidence-gap kernels
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2a_pytest.txt; tail -3 gpurun_out/r2a_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench_n1.json 2> gpurun_out/r2a_bench_err.txt; tail -c 600 gpurun_out/r2a_bench_n1.json; tail -5 gpurun_out/r2a_bench_err.txt
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2a_bench_ref.json 2>> gpurun_out/r2a_bench_err.txt
python bench_ops.py --workload all --iters 30 --json gpurun_out/r2a_bench_ops.jsonl 2>&1 | grep -v "^\[" > gpurun_out/r2a_bench_ops.txt
python bench_ops.py --workload ops2 --iters 30 --json gpurun_out/r2a_bench_ops.jsonl 2>&1 | grep -v "^\[" >> gpurun_out/r2a_bench_ops.txt
python bench_ops.py --workload pcie --iters 20 --json gpurun_out/r2a_pcie.jsonl 2>&1 | tail -2 >> gpurun_out/r2a_bench_ops.txt
cat gpurun_out/r2a_bench_ops.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2a_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-ceiling > gpurun_out/r2a_ncu_launch.log 2>&1
for kc in "pipe_kernel:c2:fused_head" "u8c3_pipe:c1:c1_point" "yuv2bgr:cvt:yuv2bgr" "layout_cn:layout:layout_cn" "sums_hwc3:sums:sums_hwc3" "u8_to_f32:dtype:u8_to_f32" "walk2:c4:cubic_walk2_base"; do
  k=${kc%%:*}; rest=${kc#*:}; c=${rest%%:*}; name=${rest#*:}
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o gpurun_out/r2_$name -f python profiles/_once.py $c > gpurun_out/r2a_ncu_$name.log 2>&1; tail -1 gpurun_out/r2a_ncu_$name.log | cut -c1-150
done
ls -la gpurun_out | tail -20
