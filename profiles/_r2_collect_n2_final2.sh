# round 2, 2-GPU trip: multi-GPU tests (NCCL + peer-memory transports of config 5, one thread alternating two GPUs), bench at N=2, PCIe ceiling at N=2
set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2f_n2_topo.txt 2>&1
python -m pytest tests/test_gpu_dist.py tests/test_gpu_host.py tests/test_gpu_multi.py -q 2>&1 | tail -5 > gpurun_out/r2f_n2_pytest.txt; cat gpurun_out/r2f_n2_pytest.txt
python -m pytest tests/test_gpu_parity.py -q -k "dropin" 2>&1 | tail -3 >> gpurun_out/r2f_n2_pytest.txt
tests/cpp/test_dropin | tail -8 >> gpurun_out/r2f_n2_pytest.txt; tail -9 gpurun_out/r2f_n2_pytest.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2f_bench_n2.json 2> gpurun_out/r2f_bench_n2.err; tail -c 1500 gpurun_out/r2f_bench_n2.json; tail -3 gpurun_out/r2f_bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench_ops.py --workload pcie --iters 20 --json gpurun_out/r2f_pcie_n2.jsonl 2>&1 | tail -2
