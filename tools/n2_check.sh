SC_SCATTER_EXCHANGE=1 python -m pytest tests/test_sharded_prover.py -m gpu -x -q -k "nccl" 2>&1 | tail -2
SC_SCATTER_EXCHANGE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2j_scale_n2.json 2> gpurun_out/r2j_scale_n2.err
tail -c 200 gpurun_out/r2j_scale_n2.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2j_scale_n2.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["gpu_launches"], d["config"]["golden_match"])
k=d["kernel_ms_per_proof"]; print({x:k[x] for x in k if "nccl" in x or "push" in x or "pack" in x})
PY
