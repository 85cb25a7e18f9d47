#!/bin/bash
# Round-end measurement set (run through gpurun from the repo root; writes into gpurun_out/final/).
#   tools/final_measure.sh 1      one GPU: tests, smoke, bench line, ncu launch list + full capture, phase stamps
#   tools/final_measure.sh N      N = 2, 4, 8: sharded check, in-kernel phase probe, bench line
set -u
N=${1:-1}
OUT=gpurun_out/final
mkdir -p $OUT
export PATH=/usr/local/cuda/bin:$PATH
if [ "$N" = "1" ]; then
  timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -4 > $OUT/pytest_gpu.txt
  timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.txt 2>&1
  timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench_n1.json 2> $OUT/bench_n1.err
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/flat512_launches.csv \
      python bench.py --steps 2 --warmup 3 --no-configs > $OUT/ncu_bench.log 2>&1
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:flat_step_kernel -s 2 -c 1 -o $OUT/fused_step \
      python tools/run_flat_step.py 512 4 > $OUT/ncu_full.log 2>&1
  timeout 120 python tools/e2e_probe.py 2>&1 | grep -v Warn | head -24 | cut -c1-400 > $OUT/e2e_probe.txt
  timeout 60 python tools/time_sim_fwd.py > $OUT/sim_fwd_one_vs_two_pass.txt 2>&1
  tail -2 $OUT/pytest_gpu.txt; cat $OUT/smoke.txt | tail -1; tail -c 300 $OUT/bench_n1.json
else
  P=$((29500 + N))
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P \
      tools/sharded_check.py $((512 * N)) > $OUT/sharded_check_n$N.log 2>&1
  grep "SHARDED_OK\|Error\|assert" $OUT/sharded_check_n$N.log | head -3
  timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P + 10)) \
      tools/fused_sharded_probe.py 2>&1 | grep -v "^\*\*\|OMP_NUM\|Warn" > $OUT/sharded_probe_n$N.txt
  grep "cold:\|warm:" $OUT/sharded_probe_n$N.txt
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P + 20)) \
      bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
  python -c "import json; d=json.loads(open('$OUT/bench_n$N.json').read().strip().splitlines()[-1]); print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['configs']['config3']['ms'], d['configs']['config3']['launch_form'])"
fi
