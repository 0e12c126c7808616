set -x
O=gpurun_out/r3e
mkdir -p $O
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29588"
timeout 600 $T bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; tail -c 300 $O/bench_n2.json; tail -3 $O/bench_n2.err
python - <<'PY'
import os
for n in (0,1):
    try: print(n, open('/sys/devices/system/node/node%d/cpulist'%n).read().strip())
    except Exception as e: print(n, e)
import glob
for p in glob.glob('/sys/bus/pci/devices/*/numa_node')[:400]:
    v=open(p).read().strip()
    if v not in ('-1',): print(p, v)
PY
