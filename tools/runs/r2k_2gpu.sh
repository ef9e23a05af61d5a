set -x
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/pytest_r2k_2gpu.log 2>&1; echo "pytest rc $?"
tail -6 gpurun_out/pytest_r2k_2gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_r2k_2gpu.json 2> gpurun_out/bench_r2k_2gpu.err; echo "bench rc $?"
tail -c 1500 gpurun_out/bench_r2k_2gpu.json
python tools/sampler_sweep.py 1,2 1000,20000 2>&1 | grep -E "default" | tee gpurun_out/sampler_sweep_r2k_2gpu.log
