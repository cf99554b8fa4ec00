#!/bin/bash
# Round-2 validation on the GPU box (through gpurun): GPU parity suite, smoke, bench line, then launch lists of the
# detector forward (256 frames) and of the training step (64 frames) with the library's kernels only.
set -o pipefail
TAG=${1:-r2a}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/${TAG}_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/${TAG}_smoke.log
timeout 400 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
tail -c 400 gpurun_out/${TAG}_bench.json
KRE='regex:^(knn|mp_|rowmlp|wgrad|dproj|tile_program|pack|chain64|conv_nodes|edge_enc|emb_split|loss|sgd|sym_|sort_|src_|finalize|node_feat|edge_feat|scan_|proposals|cc_|uf_|seg)'
if [ "$2" != "nolist" ]; then
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KRE" -c 400 --csv \
  --log-file gpurun_out/${TAG}_launches_fwd.csv python tools/prof_layer.py --what forward --frames 256 --reps 1 > gpurun_out/${TAG}_ncu_fwd.log 2>&1
echo "ncu fwd rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KRE" -c 1500 --csv \
  --log-file gpurun_out/${TAG}_launches_train.csv python tools/prof_layer.py --what train --frames 64 --reps 1 > gpurun_out/${TAG}_ncu_train.log 2>&1
echo "ncu train rc=$?"
fi
