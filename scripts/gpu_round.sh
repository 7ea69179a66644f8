#!/bin/bash
# One GPU-box session: tests, bench lines, launch list.  usage (via gpurun): bash scripts/gpu_round.sh <tag> [stages...]
# stages: tests bench ref train ncu_launch   (default: tests bench ref)
tag=${1:-r2}; shift
stages=${@:-tests bench ref}
out=gpurun_out
mkdir -p $out
for st in $stages; do
  case $st in
    tests)  timeout 1500 python -m pytest tests -x -q -m gpu -s > $out/${tag}_tests.log 2>&1; echo "tests rc=$?"; tail -5 $out/${tag}_tests.log ;;
    bench)  timeout 900 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"; tail -c 3000 $out/${tag}_bench.json; tail -5 $out/${tag}_bench.err ;;
    ref)    timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > $out/${tag}_ref.json 2> $out/${tag}_ref.err; echo "ref rc=$?"; tail -c 1500 $out/${tag}_ref.json ;;
    train)  timeout 600 python bench.py --workload train > $out/${tag}_train.json 2> $out/${tag}_train.err; echo "train rc=$?"; tail -c 2500 $out/${tag}_train.json; tail -3 $out/${tag}_train.err ;;
    traintests) timeout 900 python -m pytest tests/test_gpu_full_size.py tests/test_gpu_parity.py tests/test_gpu_training_loop.py -x -q -m gpu -k 'train or grad or trainer' > $out/${tag}_traintests.log 2>&1; echo "traintests rc=$?"; tail -4 $out/${tag}_traintests.log ;;
    sanitize)
      for tool in memcheck racecheck; do timeout 900 compute-sanitizer --tool $tool python scripts/sanitize_stream_kernels.py > $out/${tag}_sanitizer_${tool}_stream.log 2>&1; tail -2 $out/${tag}_sanitizer_${tool}_stream.log; done
      timeout 900 compute-sanitizer --tool memcheck python scripts/sanitize_chain_kernels.py > $out/${tag}_sanitizer_memcheck_chain.log 2>&1; tail -3 $out/${tag}_sanitizer_memcheck_chain.log ;;
    parity) timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -s > $out/${tag}_parity.log 2>&1; echo "parity rc=$?"; grep -i 'x1.5\|passed\|failed\|Error' $out/${tag}_parity.log | tail -8 ;;
    smoke)  timeout 600 python -c 'import __graft_entry__ as g; g.smoke()' > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $out/${tag}_smoke.log ;;
    ncu_launch)
      ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches_train.csv python bench.py --workload train --steps 3 --warmup 3 > $out/${tag}_ncu_train.log 2>&1
      ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_render.csv python bench.py --workload render --steps 3 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_render.log 2>&1 ;;
    multirank) timeout 900 python -m pytest tests/test_gpu_multirank.py -x -q -s > $out/${tag}_multirank.log 2>&1; echo "multirank rc=$?"; tail -15 $out/${tag}_multirank.log ;;
    bench2) timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $out/${tag}_bench2.json 2> $out/${tag}_bench2.err; echo "bench2 rc=$?"; tail -c 3000 $out/${tag}_bench2.json; tail -5 $out/${tag}_bench2.err ;;
    bench8) for n in 4 8; do timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 20 --warmup 5 > $out/${tag}_bench$n.json 2> $out/${tag}_bench$n.err; echo "bench$n rc=$?"; tail -c 2500 $out/${tag}_bench$n.json; done ;;
  esac
done
