P="python tools/probe.py --workload cfg4 --nfreq 8 --flo 4 --fhi 8 --reps 2"
for cfg in "1 256 4 512 64 256" "1 256 2 256 32 128" "2 512 2 256 16 128" "1 128 1 256 32 256" "2 256 4 256 64 512"; do
  set -- $cfg
  echo "cfg vx=$1 thx=$2 vy=$3 thy=$4 vz=$5 thz=$6" >> gpurun_out/sweep.txt
  FV_T3_VX=$1 FV_T3_THRX=$2 FV_T3_VY=$3 FV_T3_THRY=$4 FV_T3_VZ=$5 FV_T3_THRZ=$6 $P 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['run_s'], d['stages'])" >> gpurun_out/sweep.txt
done
