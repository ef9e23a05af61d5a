set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/pytest_r2r.log 2>&1; echo "pytest rc $?"
tail -4 gpurun_out/pytest_r2r.log
python bench.py --steps 2 --warmup 3 > gpurun_out/bench_r2r.json 2> gpurun_out/bench_r2r.err; echo "bench rc $?"
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_r2r.json 2>&1; echo "ref rc $?"
tail -c 400 gpurun_out/bench_ref_r2r.json
