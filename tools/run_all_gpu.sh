set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for w in teacher_nll student generate distill encode; do
  python bench.py --workload $w --steps 5 --warmup 3 > gpurun_out/r01e_bench_$w.json 2> gpurun_out/r01e_bench_$w.err || echo "bench $w failed"
  tail -c 600 gpurun_out/r01e_bench_$w.json
done
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01e_bench_reference.json 2>&1; tail -c 400 gpurun_out/r01e_bench_reference.json
ncu --set full --import-source on --clock-control none -k regex:k_ar_mma -c 1 -o gpurun_out/r01e_ar_full -f python tools/prof_ar.py 256x256 fp16 1 > gpurun_out/ncu_ar_full.log 2>&1; tail -1 gpurun_out/ncu_ar_full.log
ncu --set full --import-source on --clock-control none -k regex:k_bwd_gate\|k_bwd_conv -s 20 -c 2 -o gpurun_out/r01e_bwd_full -f python bench.py --workload distill --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bwd_full.log 2>&1; tail -1 gpurun_out/ncu_bwd_full.log
