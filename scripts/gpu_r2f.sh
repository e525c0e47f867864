#!/bin/bash
# Round 2, GPU call F (8 GPUs): config 5 (1 M sentences, CoNg, round-robin shards) through ONE handle over N devices (kiwi_b200_init_multi)
# at N = 1, 2, 4, 8 (strong scaling), the same job under torchrun at N = 8, and the default weak-scaling line at N = 8.
mkdir -p gpurun_out; O=gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > $O/r2f_smi.txt 2>&1
( timeout 600 python bench.py --config 5 --gpus 1 --steps 2 --warmup 1 --no-cpu ) > $O/r2f_cfg5_n1.json 2> $O/r2f_cfg5_n1.err
for n in 2 4 8; do
  ( timeout 600 python bench.py --config 5 --gpus $n --inproc --steps 2 --warmup 1 --no-cpu ) > $O/r2f_cfg5_inproc_n$n.json 2> $O/r2f_cfg5_inproc_n$n.err
done
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --config 5 --gpus 8 --steps 2 --warmup 1 --no-cpu ) > $O/r2f_cfg5_torchrun_n8.json 2> $O/r2f_cfg5_torchrun_n8.err
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu ) > $O/r2f_cfg2_torchrun_n8.json 2> $O/r2f_cfg2_torchrun_n8.err
( timeout 300 python -m pytest tests/test_gpu_round2.py -m gpu -q -k "multi_device" -p no:cacheprovider ) > $O/r2f_pytest_multi.log 2>&1
for f in $O/r2f_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("n_gpus %s value %.0f e2e %.0f ms/step %.2f order_ok %s" % (d["n_gpus"], d["value"], d["e2e"]["value"], d["ms_per_step"], d["config"].get("ordered_merge_check")))
except Exception as e:
    print("ERR", e, open(sys.argv[1].replace(".json",".err")).read()[-800:])
PY
done
tail -n 3 $O/r2f_pytest_multi.log
