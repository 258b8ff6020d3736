#!/bin/bash
# multi-GPU trip: parity of the sharded paths (torch collectives and the C ABI's own NCCL exchange), then the bench at N GPUs
n=${1:-2}; tag=${2:-mg}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 scripts/multi_gpu_check.py > gpurun_out/${tag}_check.txt 2>&1
echo "multi_gpu_check rc=$?"; grep -v "^\*\|^$\|OMP_NUM\|Setting" gpurun_out/${tag}_check.txt | tail -15
python -m pytest tests/test_multi_gpu.py -q -x -m gpu 2>&1 | tail -3
python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $n --steps 200 --warmup 5 > gpurun_out/${tag}_bench_n${n}.json 2> gpurun_out/${tag}_bench_n${n}.err
tail -3 gpurun_out/${tag}_bench_n${n}.err
python - <<PY
import json
for f in ("gpurun_out/${tag}_bench_n1.json", "gpurun_out/${tag}_bench_n${n}.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        s=d.get("sharded") or {}
        print(f, "c2 fps %.1f e2e %.1f" % (d["value"], d["e2e"]["value"]), "| sharded ms/frame", s.get("ms_per_frame"), "no-exchange", s.get("ms_per_frame_without_exchanges"), "speedup", s.get("speedup_vs_n1"), s.get("pass_ms_rank0"))
    except Exception as e:
        print(f, "FAILED", e)
PY
