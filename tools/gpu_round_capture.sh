set -x
python -m pytest tests -m gpu -q > gpurun_out/r02k_gputests.log 2>&1; tail -4 gpurun_out/r02k_gputests.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err; tail -2 gpurun_out/r02k_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r02k_bench_launches.csv python bench.py --steps 5 --warmup 3 --skip-cpu > gpurun_out/r02k_ncu_bench.log 2>&1
AIRICE_NCU_ONCE=1 ncu --set full --clock-control none --import-source on -k regex:airice_ -o gpurun_out/r02k_all python tools/ncu_target.py all 1e6 > gpurun_out/r02k_ncu_all.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:airice_solve -c 2 -o gpurun_out/r02k_solve1e7 python tools/ncu_target.py solve 1e7 > gpurun_out/r02k_ncu_solve.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
