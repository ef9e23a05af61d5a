# ncu of the FINAL sampler launch policy (second launch = last four photon steps, order formed by lap_rank_kernel)
set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"laplace|lap_rank" -c 3 -o gpurun_out/ncu_sampler_c5_r2_final -f python tools/profile_sampler.py 1000 > gpurun_out/ncu_sampler_c5_r2_final.log 2>&1
tail -2 gpurun_out/ncu_sampler_c5_r2_final.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_sampler_r2_final.csv python tools/profile_sampler.py 1000 > /dev/null 2>&1
tail -5 gpurun_out/launches_sampler_r2_final.csv
ls -la gpurun_out/ncu_sampler_c5_r2_final.ncu-rep
