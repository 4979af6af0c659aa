#!/bin/bash
# ncu launch list of ONE bench step per precision mode -> gpurun_out/<tag>_launches_bench_step_<mode>.csv and
# gpurun_out/<tag>_traffic_<mode>.json (copy both into profiles/: bench.py reads the newest profiles/r2*_traffic_<mode>.json
# for roofline.traffic).  Run on the GPU box from the build container:
#     gpurun -- "Y2_HEAD=$(git rev-parse --short HEAD) bash tools/profile_step.sh r2a"
# ncu durations are cold-cache and serialised: only DRAM bytes and the kernels' SHARE of the step are taken from them.
set -u
TAG=${1:-r2}
mkdir -p gpurun_out
for MODE in bf16 bf16x3; do
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
      --log-file gpurun_out/${TAG}_ncu_raw_${MODE}.csv \
      python bench.py --steps 3 --warmup 3 --no-graph --single-mode --precision ${MODE} --no-cpu-baseline --sustain-seconds 0 \
      > gpurun_out/${TAG}_ncu_bench_${MODE}.log 2>&1
  python tools/ncu_launch_list.py gpurun_out/${TAG}_ncu_raw_${MODE}.csv gpurun_out/${TAG} ${MODE}
  rm -f gpurun_out/${TAG}_ncu_raw_${MODE}.csv
done
