# Round capture (run under gpurun on one B200): GPU tests, bench (both arms), launch list of the bench, ncu --set full of
# every kernel.  gpurun_out/ is limited to 64 MiB, so each report is condensed ON THE BOX by tools/ncu_summary.py
# (per-launch metrics + opcode mix / hottest instructions per kernel) and then deleted; the summaries are what
# profiles/ keeps (tools/kernel_facts.py turns them into profiles/kernel_facts.json).
# usage: bash tools/gpu_round_capture.sh <prefix> [part ...]   parts: tests bench all solve lookup multi (default: all of them)
set -x
P=${1:-r02f}; shift
PARTS=${@:-tests bench all solve lookup multi}
mkdir -p gpurun_out
cap() {   # cap <name> <kernel regex> <count or 0> <target command...>
  name=$1; k=$2; c=$3; shift 3
  if [ "$c" = 0 ]; then cnt=""; else cnt="-c $c"; fi
  ncu --set full --clock-control none --import-source on -k regex:"$k" $cnt -o gpurun_out/${P}_$name "$@" > gpurun_out/${P}_ncu_$name.log 2>&1
  python tools/ncu_summary.py gpurun_out/${P}_$name.ncu-rep gpurun_out/${P}_$name
  rm -f gpurun_out/${P}_$name.ncu-rep
}
for part in $PARTS; do
  case $part in
    tests) python -m pytest tests -m gpu -q > gpurun_out/${P}_gputests.log 2>&1; tail -4 gpurun_out/${P}_gputests.log;;
    bench) python bench.py --steps 20 --warmup 3 > gpurun_out/${P}_bench.json 2> gpurun_out/${P}_bench.err; tail -2 gpurun_out/${P}_bench.err
           python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${P}_bench_reference.json 2> gpurun_out/${P}_bench_reference.err
           ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${P}_bench_launches.csv python bench.py --steps 5 --warmup 3 --skip-cpu > gpurun_out/${P}_ncu_bench.log 2>&1;;
    all) AIRICE_NCU_ONCE=1 cap all airice_ 0 python tools/ncu_target.py all 1e6;;
    solve) cap solve1e7 airice_solve 2 python tools/ncu_target.py solve 1e7;;
    lookup) cap lookup1e7 airice_lookup 1 python tools/ncu_target.py lookup 1e7;;
    multi) cap multi64 'table_multi|row_' 4 python tools/multi64_target.py;;
  esac
done
du -sh gpurun_out; ls gpurun_out | head -80
