#!/bin/bash
# Closes the open verification item of DESIGN.md 6: the 17-stage Haar ALL cascade (BASELINE configs[2]) must print the same
# stage list and the same `neg_consumed` sequence for every world size.  Usage (on a box with >= 8 GPUs):
#   bash tools/verify_world_sizes.sh "1 2 4 8"
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ref=""
rc=0
for w in ${1:-1 2 8}; do
  if [ "$w" = 1 ]; then cmd="python tools/full_cascade.py 20 5000 10000 48"
  else cmd="python -m torch.distributed.run --nnodes=1 --nproc-per-node $w --master-addr 127.0.0.1 --master-port $((29600 + w)) tools/full_cascade.py 20 5000 10000 48"; fi
  line=$($cmd 2>&1 | grep '"neg_consumed"' | tail -1)
  echo "$line" > gpurun_out/world_$w.json
  key=$(python - "$line" <<'PY'
import json, sys
d = json.loads(sys.argv[1])
print(json.dumps([d["weak"], d["neg_consumed"]]))
PY
)
  t=$(python -c "import json,sys; d=json.loads(sys.argv[1]); print('%.2f s (mining %.2f s)' % (d['total_s'], d['mining_s']))" "$line")
  echo "world=$w: $t  $key"
  if [ -z "$ref" ]; then ref="$key"; elif [ "$key" != "$ref" ]; then echo "  MISMATCH against the first world size"; rc=1; fi
done
[ $rc = 0 ] && echo "all world sizes agree"
exit $rc
