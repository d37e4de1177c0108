#!/bin/bash
# Round evidence on one B200 (run through gpurun): GPU parity suite, the N=1 bench line, the reference arm,
# the other BASELINE configs, the ncu launch list of the bench command and full captures of the top kernels.
# Everything lands in gpurun_out/ev/ (copy what is to be judged into profiles/).
set -u
O=gpurun_out/ev; mkdir -p $O
(time timeout 400 python -m pytest tests -m gpu -x -q) > $O/pytest_gpu.log 2>&1; tail -2 $O/pytest_gpu.log
python -c 'import __graft_entry__ as g; g.smoke()' > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 600 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; tail -c 400 $O/bench_n1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 600 python bench_configs.py > $O/other_configs.jsonl 2> $O/other_configs.err; tail -3 $O/other_configs.err
if [ -z "${SKIP_STEP_NCU:-}" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-conv-vjp --no-graph > $O/launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:(conv_.._kernel|strip_.*_kernel)' -s 10 -c 8 -f -o $O/conv_step \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-conv-vjp --no-graph > $O/conv_step.log 2>&1
fi
# ConvVjpBench shape: launch 3 = forward, 4-7 = dKrn, 8 = dInp
timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:(conv_.._kernel)' -s 3 -c 6 -f -o $O/conv_vjp \
  python tools/convvjp_profile.py 2 > $O/conv_vjp.log 2>&1
ls -la $O
