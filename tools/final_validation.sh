set -x
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 300 python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; tail -1 gpurun_out/bench_default.log | cut -c1-300
timeout 200 python bench.py --criterion kl --no-cpu-baseline --no-feeders --no-contraction > gpurun_out/bench_kl.log 2>&1; tail -1 gpurun_out/bench_kl.log | cut -c1-200
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.log 2>&1; tail -1 gpurun_out/bench_reference.log | cut -c1-300
timeout 300 python tools/mode_sweep.py > gpurun_out/mode_sweep.txt 2>&1; grep -c us gpurun_out/mode_sweep.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-contraction --no-feeders --no-graph > gpurun_out/ncu_launches_final.log 2>&1; tail -2 gpurun_out/ncu_launches_final.log | cut -c1-200
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dsgfd_kl_col -c 1 -o gpurun_out/prof_kl_col_final -f python bench.py --criterion kl --no-cpu-baseline --no-feeders --no-contraction --no-e2e --steps 2 --warmup 1 > gpurun_out/ncu_col_final.log 2>&1; tail -1 gpurun_out/ncu_col_final.log | cut -c1-120
