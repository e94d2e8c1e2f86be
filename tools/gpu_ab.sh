#!/bin/bash
# A/B of RIG_VARIANT settings on one workload: bench line + light ncu metrics of the expansion kernels.
# usage: tools/gpu_ab.sh <tag> <workload> <variant> [<variant> ...]
tag=$1; wl=$2; shift 2
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_lookup_miss.sum,lts__t_sectors_lookup_hit.sum,lts__t_tag_requests.avg.pct_of_peak_sustained_elapsed,l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active
for v in "$@"; do
  RIG_VARIANT=$v python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu $BENCH_EXTRA > gpurun_out/${tag}_${wl}_v$v.json 2>> gpurun_out/${tag}.err
  RIG_VARIANT=$v ncu --metrics $M --clock-control none -k regex:phi_ -s 4 -c 2 --csv --log-file gpurun_out/${tag}_${wl}_v$v.ncu.csv python bench.py --workload $wl --steps 1 --warmup 3 --no-cpu $BENCH_EXTRA > /dev/null 2>> gpurun_out/${tag}.err
done
