#!/bin/bash
# 8-GPU trip: N=1 record, then the bench at N GPUs (frame replicas headline + the `sharded` record) and the multi-GPU parity check
# (scripts/pcie_probe.py, the host I/O probe, is run separately: profiles/r2_pcie_probe.txt)
n=${1:-8}; tag=${2:-mg8}; mids=${3:-"2 4"}
mkdir -p gpurun_out
python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
for k in $mids $n; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $k --master-addr 127.0.0.1 --master-port 2961$k bench.py --gpus $k --steps 200 --warmup 5 > gpurun_out/${tag}_bench_n${k}.json 2> gpurun_out/${tag}_bench_n${k}.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29641 scripts/multi_gpu_check.py > gpurun_out/${tag}_check.txt 2>&1
echo "multi_gpu_check rc=$?"; grep -v "^\*\|^$\|OMP_NUM\|Setting" gpurun_out/${tag}_check.txt | tail -12
python - <<PY
import json
for k in (1, 2, 4, $n):
    f = "gpurun_out/${tag}_bench_n%d.json" % k
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        s=d.get("sharded") or {}
        print(k, "c2 fps %.1f e2e %.1f" % (d["value"], d["e2e"]["value"]), "| sharded ms/frame %.3f" % s.get("ms_per_frame"), "no-exchange", s.get("ms_per_frame_without_exchanges"), "speedup", s.get("speedup_vs_n1"), {a: round(b, 3) for a, b in s.get("pass_ms_rank0").items()})
    except Exception as e:
        print(f, "FAILED", e)
PY
tail -2 gpurun_out/${tag}_bench_n${n}.err
