set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:glynn_plain -s 1 -c 1 -o gpurun_out/ncu_plain40_r2 -f python tools/profile_plain.py 40 512 3 > gpurun_out/ncu_plain40_r2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:glynn_reppow -s 1 -c 1 -o gpurun_out/ncu_pow_c4_r2 -f python tools/profile_repeated.py c4 > gpurun_out/ncu_pow_c4_r2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:glynn_reppow -s 1 -c 1 -o gpurun_out/ncu_pow_n60_r2 -f python tools/profile_repeated.py n60 > gpurun_out/ncu_pow_n60_r2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:laplace -c 2 -o gpurun_out/ncu_sampler_c5_r2 -f python tools/profile_sampler.py 1000 > gpurun_out/ncu_sampler_c5_r2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench_r2.csv python bench.py --steps 2 --warmup 3 --no-configs > gpurun_out/bench_under_ncu_r2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_sampler_r2.csv python tools/profile_sampler.py 1000 > /dev/null 2>&1
tail -2 gpurun_out/ncu_plain40_r2.log gpurun_out/ncu_pow_c4_r2.log gpurun_out/ncu_pow_n60_r2.log gpurun_out/ncu_sampler_c5_r2.log
ls -la gpurun_out | tail -12
