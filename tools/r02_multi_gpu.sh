#!/usr/bin/env bash
# Round-2 multi-GPU evidence on one 8-GPU box: C5 (row bands, NCCL halo refreshes) at 2/4/8 GPUs, C4 routing and C3 strong /
# weak scaling at 8, the NCCL band parity test.  Usage: gpurun --gpus 8 -- bash tools/r02_multi_gpu.sh
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/r02m_topo.txt 2>&1; lscpu | grep -i "numa\|model name\|^CPU(s)" >> gpurun_out/r02m_topo.txt
timeout 300 python -m pytest tests/test_gpu_band.py -x -q -k nccl 2>&1 | tail -2
python bench.py --config c5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02m_c5_n1.json 2> gpurun_out/r02m_c5_n1.err
for n in 2 4 8; do
  timeout 300 $TR --nproc-per-node $n --master-port $((29600 + n)) bench.py --config c5 --gpus $n --steps 10 --warmup 3 > gpurun_out/r02m_c5_n$n.json 2> gpurun_out/r02m_c5_n$n.err
done
timeout 300 $TR --nproc-per-node 8 --master-port 29621 bench.py --config c5 --gpus 8 --halo 32 --steps 10 --warmup 3 > gpurun_out/r02m_c5_n8_halo32.json 2> gpurun_out/r02m_c5_n8_halo32.err
timeout 300 $TR --nproc-per-node 8 --master-port 29622 bench.py --config c4 --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02m_c4_n8.json 2> gpurun_out/r02m_c4_n8.err
timeout 300 $TR --nproc-per-node 8 --master-port 29623 bench.py --gpus 8 --scaling strong --steps 10 --warmup 3 > gpurun_out/r02m_c3_strong_n8.json 2> gpurun_out/r02m_c3_strong_n8.err
timeout 300 $TR --nproc-per-node 8 --master-port 29624 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02m_c3_weak_n8.json 2> gpurun_out/r02m_c3_weak_n8.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02m_*.json")):
    try:
        d = json.load(open(f))
        print(f.split("/")[-1], d["n_gpus"], "ms", round(d["ms_per_step"], 3), d["unit"], round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1),
              "u8", round(d.get("e2e_u8", {}).get("value", 0), 1))
    except Exception as e:
        print(f, "unreadable", e)
PY
