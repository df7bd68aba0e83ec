#!/bin/bash
# torchrun --no-python tools/rank0_ncu.sh <out.csv> <python args...>: rank 0 runs under `ncu --metrics gpu__time_duration.sum`
# (a launch list, no replay), the other ranks run plainly.  Numbers printed by such a run are never bench values.
out="$1"; shift
if [ "${LOCAL_RANK:-0}" = "0" ]; then
  exec ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file "$out" python "$@"
else
  exec python "$@"
fi
