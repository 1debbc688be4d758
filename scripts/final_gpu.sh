set -x
mkdir -p gpurun_out/z1
(timeout 900 python -m pytest tests -m gpu -q > gpurun_out/z1/gpu_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/z1/gpu_tests.log)
(timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/z1/smoke.log 2>&1)
for wl in cfg1 cfg3 cfg4 cfg5; do timeout 300 python bench.py --workload $wl --no-cpu-baseline --no-strong > gpurun_out/z1/bench_$wl.json 2> gpurun_out/z1/bench_$wl.err; done
timeout 400 python bench.py > gpurun_out/z1/bench_cfg2.json 2> gpurun_out/z1/bench_cfg2.err
timeout 300 python bench.py --padding masked --no-cpu-baseline --no-strong > gpurun_out/z1/bench_cfg2_masked.json 2> gpurun_out/z1/bench_cfg2_masked.err
timeout 300 python bench.py --engine bf16 --no-cpu-baseline --no-strong > gpurun_out/z1/bench_cfg2_bf16.json 2> gpurun_out/z1/bench_cfg2_bf16.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/z1/bench_cfg2_reference.json 2> gpurun_out/z1/bench_cfg2_reference.err
# launch list of an eager cfg2 step
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/z1/launches_cfg2.csv python bench.py --workload cfg2 --steps 2 --warmup 1 --no-graphs --no-cpu-baseline --no-strong --no-sustained > gpurun_out/z1/ncu_launches.log 2>&1
python scripts/agg_launches.py gpurun_out/z1/launches_cfg2.csv > gpurun_out/z1/launches_cfg2.summary.txt
# ncu --set full: one layer forward + backward of each workload (skip the warm-up step's launches)
for wl in cfg2 cfg3; do
  timeout 600 ncu --set full --clock-control none -k regex:'k_tc_|k_attn|k_pair_reduce|k_xtg_reduce' --launch-skip 60 -c 24 -o /tmp/full_$wl python bench.py --workload $wl --steps 1 --warmup 1 --no-graphs --no-cpu-baseline --no-strong --no-sustained > gpurun_out/z1/ncu_full_$wl.log 2>&1
  ncu -i /tmp/full_$wl.ncu-rep --page raw --csv > gpurun_out/z1/full_$wl.raw.csv 2>/dev/null
done
ls -la gpurun_out/z1
tail -3 gpurun_out/z1/gpu_tests.log; tail -2 gpurun_out/z1/smoke.log
