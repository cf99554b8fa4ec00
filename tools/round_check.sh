#!/bin/bash
# Round-end validation on the GPU box (run through gpurun): GPU parity suite, smoke, bench line, row-f3 profile; with
# a second argument "ncu" also one `ncu --set full` capture of the message kernel at the bench size (after the same
# command exited 0 without ncu).
set -o pipefail
TAG=${1:-s5}
mkdir -p gpurun_out
timeout 500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/${TAG}_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/${TAG}_smoke.log
timeout 300 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
tail -c 700 gpurun_out/${TAG}_bench.json
timeout 200 python tools/prof_accumulate.py 2>&1 | tail -1 | tee gpurun_out/accumulate_r1.json
if [ "$2" = "ncu" ]; then
  timeout 120 python tools/prof_layer.py --what edges --frames 256 --reps 3 2>&1 | tail -1 && \
    timeout 300 ncu --set full --clock-control none --import-source on -k regex:mp_edge_tc -s 1 -c 1 -f \
      -o gpurun_out/prof_r1_mp_v2 python tools/prof_layer.py --what edges --frames 256 --reps 1 > gpurun_out/ncu14.log 2>&1
  echo "ncu rc=$?"
fi
