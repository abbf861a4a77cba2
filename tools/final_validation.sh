# Round-end validation on one B200 (run through gpurun): tests, smoke, the driver's two bench arms, the per-kernel tools,
# the launch list and the full ncu captures of the roofline kernels.  Outputs land in gpurun_out/; the summaries are
# copied to profiles/.
set -x
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; tail -1 gpurun_out/bench_1gpu.json | cut -c1-300
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -1 gpurun_out/bench_reference.json | cut -c1-300
timeout 200 python tools/kl_perf.py > gpurun_out/kl_perf.txt 2>&1
timeout 200 python tools/kl_perf.py --images 4 >> gpurun_out/kl_perf.txt 2>&1
timeout 200 python tools/kernel_sweep.py > gpurun_out/kernel_sweep_mse.txt 2>&1
timeout 300 python tools/mode_sweep.py > gpurun_out/mode_sweep.txt 2>&1
timeout 200 python tools/host_breakdown.py > gpurun_out/host_breakdown.txt 2>&1
FLAGS="--steps 2 --warmup 1 --no-cpu-baseline --no-feeders --no-contraction --no-e2e --no-configs --no-train-step --no-kl --no-graph"
timeout 300 python bench.py $FLAGS > gpurun_out/bench_short.log 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_step_eager.csv python bench.py $FLAGS > gpurun_out/ncu_launches.log 2>&1; tail -1 gpurun_out/ncu_launches.log | cut -c1-200
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dsgfd_mse_nchw -c 1 -f -o gpurun_out/prof_mse_final python bench.py $FLAGS > gpurun_out/ncu_mse.log 2>&1; tail -1 gpurun_out/ncu_mse.log | cut -c1-120
timeout 300 python bench.py --criterion kl $FLAGS > gpurun_out/bench_short_kl.log 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_kl_step_eager.csv python bench.py --criterion kl $FLAGS > gpurun_out/ncu_launches_kl.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dsgfd_kl_stream -c 1 -f -o gpurun_out/prof_kl_final python bench.py --criterion kl $FLAGS > gpurun_out/ncu_kl.log 2>&1; tail -1 gpurun_out/ncu_kl.log | cut -c1-120
