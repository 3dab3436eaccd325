# round 2, last single-GPU trip at HEAD (third session): tests, smoke, bench line (+ reference arm), per-operator table (device-time samples), PCIe ceiling,
# launch list, ncu --set full captures of the kernels of this session and of the headline (summaries: profiles/_ncu_summary.py)
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2f_pytest.txt; tail -3 gpurun_out/r2f_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_err.txt; tail -c 700 gpurun_out/r2f_bench_n1.json; tail -3 gpurun_out/r2f_bench_err.txt
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2f_bench_ref.json 2>> gpurun_out/r2f_bench_err.txt; tail -c 400 gpurun_out/r2f_bench_ref.json
rm -f gpurun_out/r2f_bench_ops.jsonl
python bench_ops.py --workload all --iters 30 --json gpurun_out/r2f_bench_ops.jsonl 2>&1 | grep -v "^\[" > gpurun_out/r2f_bench_ops.txt
python bench_ops.py --workload ops2 --iters 30 --json gpurun_out/r2f_bench_ops.jsonl 2>&1 | grep -v "^\[" >> gpurun_out/r2f_bench_ops.txt
python bench_ops.py --workload pcie --iters 20 --json gpurun_out/r2f_pcie.jsonl 2>&1 | tail -2 >> gpurun_out/r2f_bench_ops.txt
cat gpurun_out/r2f_bench_ops.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-ceiling > gpurun_out/r2f_ncu_launch.log 2>&1
for kc in ${NCU_CASES:-"pipe_kernel:c2:fused_head" "pipe_kernel:c2p:fused_p2048_maps" "linear3_period:lin720:lin_period_head" "pack_kernel:c3u8:warp_pack_head" "cubic3_period:c4:cubic_period_head" "u8c3_pipe:c1:c1_point_head"}; do
  k=${kc%%:*}; rest=${kc#*:}; c=${rest%%:*}; name=${rest#*:}
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o gpurun_out/r2f_$name -f python profiles/_once.py $c > gpurun_out/r2f_ncu_$name.log 2>&1; tail -1 gpurun_out/r2f_ncu_$name.log | cut -c1-150
done
ls -la gpurun_out | grep r2f
