# Final measurement suite of a round (1 GPU): tests + smoke with the product library, every bench line, launch list.
set -x
O=gpurun_out/${1:-z1}
mkdir -p $O
(timeout 900 python -m pytest tests -m gpu -q > $O/gpu_tests.log 2>&1; echo "tests rc=$?" >> $O/gpu_tests.log)
(timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > $O/smoke.log 2>&1)
for wl in cfg1 cfg3 cfg4 cfg5; do timeout 300 python bench.py --workload $wl --no-cpu-baseline --no-strong > $O/bench_$wl.json 2> $O/bench_$wl.err; done
timeout 400 python bench.py > $O/bench_cfg2.json 2> $O/bench_cfg2.err
timeout 300 python bench.py --padding masked --no-cpu-baseline --no-strong > $O/bench_cfg2_masked.json 2> $O/bench_cfg2_masked.err
timeout 300 python bench.py --engine bf16 --no-cpu-baseline --no-strong > $O/bench_cfg2_bf16.json 2> $O/bench_cfg2_bf16.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_cfg2_reference.json 2> $O/bench_cfg2_reference.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_cfg2.csv python bench.py --workload cfg2 --steps 2 --warmup 1 --no-graphs --no-cpu-baseline --no-strong --no-sustained > $O/ncu_launches.log 2>&1
python scripts/agg_launches.py $O/launches_cfg2.csv > $O/launches_cfg2.summary.txt
tail -3 $O/gpu_tests.log; tail -2 $O/smoke.log
