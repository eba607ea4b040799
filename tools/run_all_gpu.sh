#!/bin/bash
# One GPU call that refreshes the measured evidence: -m gpu tests, smoke, every bench.py workload, the reference arm, the launch
# list of the default bench and ncu --set full captures of the fused and generation kernels.  Usage: bash tools/run_all_gpu.sh <tag>
tag=${1:-r01j}
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
for w in teacher_nll student generate distill encode; do
  python bench.py --workload $w --steps 5 --warmup 3 > gpurun_out/${tag}_bench_$w.json 2> gpurun_out/${tag}_bench_$w.err || echo "bench $w failed"
  python -c "import json; d=json.load(open('gpurun_out/${tag}_bench_$w.json')); print('$w', '%.4g' % d['value'], 'ms %.3f' % d['ms_per_step'], 'e2e %.4g' % d['e2e']['value'], 'frac %.3f' % d['roofline']['frac'], d['roofline']['kernel_ms'], d['clocks'])"
done
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>&1; tail -c 300 gpurun_out/${tag}_bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_bench_teacher.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_launch.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_ar_mma -c 1 -o gpurun_out/${tag}_ar_full -f python tools/prof_ar.py 256x256 fp16 1 > gpurun_out/${tag}_ncu_ar.log 2>&1; tail -1 gpurun_out/${tag}_ncu_ar.log
