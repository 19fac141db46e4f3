cd $GRAFT_REPO_ROOT
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1f.json 2> gpurun_out/bench_r1f.log; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_r1f.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['extra']['phase_ms'], d['extra']['us_per_merge']); r=d['roofline']; print(r['kernel'], r['achieved'], r['frac'], r['avg_launch_us'], r['resident_split'], r['per_launch_check']); print(d['extra']['encode']); print(d['cpu_baseline']['value'], d['gpu_launches'])"
# launch list of the SAME bench command, per-launch kernels (the resident kernel cannot run under a profiler: it needs the host)
SWB_NO_PERSISTENT=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench_r1.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
# late merges (launches 5000..5300 of one step)
SWB_NO_PERSISTENT=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:merge_rows -s 5000 -c 300 --csv --log-file gpurun_out/launches_late_r1.csv python scripts/profile_step.py config2_1GB 1 5400 > gpurun_out/ncu_late.log 2>&1; echo "ncu late rc=$?"
# full captures: merge_rows (merges 3001-3002), tokenizer, encoder
SWB_NO_PERSISTENT=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:merge_rows -s 3000 -c 2 -o gpurun_out/prof_merge_rows_r1c python scripts/profile_step.py config2_1GB 1 3100 > gpurun_out/ncu_full1.log 2>&1; echo "ncu full merge rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"wt_tokenize|enc_words|enc_gather|count_rows" -c 4 -o gpurun_out/prof_load_enc_r1c python scripts/profile_encode.py > gpurun_out/ncu_full2.log 2>&1; echo "ncu full load/enc rc=$?"
ls -la gpurun_out/*.ncu-rep
