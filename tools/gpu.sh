#!/bin/bash
# One parameterised lease script (run on a B200 box through gpurun; everything it writes goes to gpurun_out/):
#   tools/gpu.sh build                 compile libkvq.so on the box (normally it travels prebuilt)
#   tools/gpu.sh test [pytest args]    pytest -m gpu (+ smoke)
#   tools/gpu.sh bench [bench args]    bench.py, both arms
#   tools/gpu.sh list                  ncu launch list of a short bench run
#   tools/gpu.sh full REGEX [skip] [count]   ncu --set full of the kernels matching REGEX inside a short bench run
#   tools/gpu.sh multi N               NCCL parity check + bench.py on N GPUs
# Several stages can be chained: tools/gpu.sh test -- bench -- list
set -u
mkdir -p gpurun_out
SHORT="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-cublas --no-parity --no-side --no-refcuda"

stage() {
  local what=$1; shift
  case "$what" in
    build)
      python __graft_entry__.py > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; return 1; } ;;
    test)
      nvidia-smi > gpurun_out/gpu.txt 2>&1
      timeout 1500 python -m pytest tests -m gpu -q --timeout 900 "$@" > gpurun_out/pytest_gpu.log 2>&1
      echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
      timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/smoke.log ;;
    bench)
      timeout 900 python bench.py "$@" > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
      timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err
      echo "bench ref rc=$?"; cut -c1-300 gpurun_out/bench_ref.log ;;
    list)
      $SHORT > gpurun_out/bench_plain.log 2>&1 &&
      ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $SHORT > gpurun_out/ncu_list.log 2>&1
      echo "ncu list rc=$?" ;;
    full)
      local regex=$1 skip=${2:-0} count=${3:-4} name=${4:-kernels}
      $SHORT > gpurun_out/bench_plain2.log 2>&1 &&
      ncu --set full --clock-control none --import-source on -k regex:"$regex" -s "$skip" -c "$count" -f -o gpurun_out/$name $SHORT > gpurun_out/ncu_full_$name.log 2>&1
      echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full_$name.log
      ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv 2>/dev/null ;;
    multi)
      local n=$1
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port 29541 \
        tests/dist_gpu_check.py > gpurun_out/dist_check_$n.log 2> gpurun_out/dist_check_$n.err; echo "dist check rc=$?"; tail -12 gpurun_out/dist_check_$n.log
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port 29542 \
        bench.py --gpus "$n" --steps 10 --warmup 3 > gpurun_out/bench_n$n.log 2> gpurun_out/bench_n$n.err; echo "bench n=$n rc=$?"; tail -3 gpurun_out/bench_n$n.err ;;
    *) echo "unknown stage $what"; return 2 ;;
  esac
}

args=()
for a in "$@"; do
  if [ "$a" = "--" ]; then stage "${args[@]}"; args=(); else args+=("$a"); fi
done
[ ${#args[@]} -gt 0 ] && stage "${args[@]}"
exit 0
