set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_d.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu_d.log
python bench.py > gpurun_out/r02_bench_n1_d.json 2> gpurun_out/r02_bench_n1_d.err
python bench.py --impl reference > gpurun_out/r02_bench_ref_n1_d.json 2> gpurun_out/r02_bench_ref_n1_d.err
for spec in "jacobi_2d S" "jacobi_2d L" "jacobi_2d M" "fdtd_2d M" "fdtd_2d L"; do
  set -- $spec
  ncu --set full --clock-control none --import-source on -k regex:regtile -c 1 -f -o gpurun_out/r02_prof_$1_$2 python tools/ncu_one.py $1 $2 2 > gpurun_out/r02_ncu_$1_$2.log 2>&1
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r02_bench_under_ncu.log 2>&1
ls -la gpurun_out | tail -12
