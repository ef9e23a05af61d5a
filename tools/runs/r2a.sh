set -x
mkdir -p gpurun_out
nvidia-smi -L
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2a.log 2>&1; echo "smoke rc $?"
python -m pytest tests -m gpu -q --maxfail=30 -x --timeout=900 > gpurun_out/pytest_r2a.log 2>&1; echo "pytest rc $?"
tail -5 gpurun_out/pytest_r2a.log
python tools/rep_cases.py > gpurun_out/rep_cases_r2a.log 2>&1
PQB200_NO_DD_FALLBACK=1 python tools/profile_sampler.py 1000 > gpurun_out/sampler_r2a.log 2>&1
python tools/profile_sampler.py 125 >> gpurun_out/sampler_r2a.log 2>&1
python tools/profile_sampler.py 20000 >> gpurun_out/sampler_r2a.log 2>&1
python bench.py --steps 2 --warmup 3 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; echo "bench rc $?"
ncu --set full --clock-control none --import-source on -k regex:glynn_reppow -s 1 -c 1 -o gpurun_out/ncu_pow_c4_r2a -f python tools/profile_repeated.py > gpurun_out/ncu_pow.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:laplace -s 5 -c 1 -o gpurun_out/ncu_sampler_c5_r2a -f python tools/profile_sampler.py 1000 > gpurun_out/ncu_sampler.log 2>&1
ls -la gpurun_out
