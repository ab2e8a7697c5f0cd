# Final measurements of a round on one B200 (run under gpurun): GPU tests, smoke, both bench arms, the ncu launch list
# of the bench command and full captures of the two kernels whose summaries profiles/ carries.
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
python bench.py --impl reference > gpurun_out/r02_bench_ref_n1.json 2> gpurun_out/r02_bench_ref_n1.err
ncu --set full --clock-control none --import-source on -k regex:jacobi2d_march -s 1 -c 1 -f -o gpurun_out/r02_prof_jacobi_2d_weak_late python tools/ncu_one.py jacobi_2d weak 1 > gpurun_out/r02_ncu_weak_late.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:heat3d_regtile -s 2 -c 1 -f -o gpurun_out/r02_prof_heat_3d_L_late python tools/ncu_heat_L.py > gpurun_out/r02_ncu_heat_late.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-harness > gpurun_out/r02_bench_under_ncu.log 2>&1
ls -la gpurun_out | tail -12
