set -x
cd /root/repo
timeout 600 python bench.py > gpurun_out/r02_bench_3.json 2> gpurun_out/r02_bench_3.err
tail -c 600 gpurun_out/r02_bench_3.json
timeout 600 python bench.py --workload c4 > gpurun_out/r02_bench_c4_3.json 2> gpurun_out/r02_bench_c4_3.err
tail -c 300 gpurun_out/r02_bench_c4_3.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_final.csv python bench.py --profile --steps 6 --warmup 3 > gpurun_out/r02_ncu7.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spectro_reg256 -s 6 -c 1 -o gpurun_out/r02_spectro_v7n_tg2 -f python bench.py --profile --steps 4 --warmup 3 > gpurun_out/r02_ncu8.log 2>&1
ncu --set full --clock-control none -k regex:extract2 -s 6 -c 1 -o gpurun_out/r02_extract2 -f python bench.py --profile --steps 4 --warmup 3 > gpurun_out/r02_ncu9.log 2>&1
ncu --set full --clock-control none -k regex:probe_lean -s 6 -c 1 -o gpurun_out/r02_probe_lean -f python bench.py --profile --steps 4 --warmup 3 > gpurun_out/r02_ncu10.log 2>&1
python tools/bench_configs.py --steps 20 > gpurun_out/r02_bench_configs_final.jsonl 2> gpurun_out/r02_bench_configs_final.err
tail -3 gpurun_out/r02_bench_configs_final.err
ls -la gpurun_out/*.ncu-rep | tail -4
