"""Print the host/GPU topology a multi-GPU rank sees (diagnostics for the e2e result copies)."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from or_gym_inventory_b200.sharding import _gpu_numa_node
print("cpus available:", len(os.sched_getaffinity(0)), "of", os.cpu_count())
for d in range(torch.cuda.device_count()):
    print("gpu", d, "numa", _gpu_numa_node(d))
for cmd in (["nvidia-smi", "topo", "-m"], ["sh", "-c", "ls /sys/devices/system/node | head -20; cat /sys/devices/system/node/node*/cpulist 2>/dev/null | head -8"]):
    try:
        print(subprocess.run(cmd, capture_output=True, text=True, timeout=30).stdout[-3000:])
    except Exception as e:  # noqa: BLE001
        print(cmd, e)
