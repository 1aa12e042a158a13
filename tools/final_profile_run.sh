set -x
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/z_smoke.log 2>&1; echo smoke=$?
python bench.py > gpurun_out/z_bench_n1.json 2> gpurun_out/z_bench_n1.err; echo bench=$?
python bench.py --impl reference > gpurun_out/z_bench_reference.json 2> gpurun_out/z_bench_reference.err; echo ref=$?
for w in ppo_allegro_16384x16 ppo_humanoid_4096x32; do python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-kernel-sweep --no-recipes > gpurun_out/z_$w.json 2>/dev/null; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/z_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-kernel-sweep --no-recipes > gpurun_out/z_ncu.log 2>&1; echo list=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/z_launches_yaml.csv python bench.py --workload ppo_ant_4096x32_yaml --steps 1 --warmup 3 --no-cpu-baseline --no-kernel-sweep --no-recipes > gpurun_out/z_ncu_yaml.log 2>&1; echo listy=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/z_launches_allegro.csv python bench.py --workload ppo_allegro_16384x16 --steps 1 --warmup 3 --no-cpu-baseline --no-kernel-sweep --no-recipes > gpurun_out/z_ncu_allegro.log 2>&1; echo lista=$?
ncu --set full --clock-control none --import-source on -k 'regex:tc_|heads_fused|reduce_clip' --launch-skip 90 --launch-count 9 -f -o gpurun_out/z_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph --no-kernel-sweep --no-recipes > gpurun_out/z_full.log 2>&1; echo full=$?
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
