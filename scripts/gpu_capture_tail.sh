LIGHT="--steps 3 --warmup 3 --cpu-budget 0 --learner-steps 6 --fp32-steps 0 --sustained-s 0"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r02z.csv python bench.py $LIGHT > gpurun_out/ncu_bench_r02z.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:learner_tail -s 4 -c 1 -o gpurun_out/prof_tail_r02z -f python bench.py $LIGHT > gpurun_out/ncu_tail_r02z.log 2>&1
tail -1 gpurun_out/ncu_tail_r02z.log | cut -c1-120
