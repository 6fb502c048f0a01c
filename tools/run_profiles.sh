set -x
cd /root/repo
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r02_bench_v8.json 2> gpurun_out/r02_bench_v8.err
tail -c 600 gpurun_out/r02_bench_v8.json
timeout 600 python bench.py --workload c4 > gpurun_out/r02_bench_c4_v8.json 2> gpurun_out/r02_bench_c4_v8.err
tail -c 300 gpurun_out/r02_bench_c4_v8.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_v8.csv python bench.py --profile --steps 6 --warmup 3 > gpurun_out/r02_ncu_v8_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spectro_reg256 -s 6 -c 1 -o gpurun_out/r02_spectro_v8 -f python bench.py --profile --steps 4 --warmup 3 > gpurun_out/r02_ncu_v8_b.log 2>&1
python tools/bench_configs.py --steps 20 > gpurun_out/r02_bench_configs_v8.jsonl 2> gpurun_out/r02_bench_configs_v8.err
tail -3 gpurun_out/r02_bench_configs_v8.err
ls -la gpurun_out/*.ncu-rep | tail -4
