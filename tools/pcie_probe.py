"""Developer probe: where the GPU hangs (NUMA node), what this process may run on, and pinned host -> device bandwidth
with the allocating thread bound to each NUMA node in turn."""
import glob, os, subprocess, time
import torch

print("affinity:", sorted(os.sched_getaffinity(0)))
for n in sorted(glob.glob("/sys/devices/system/node/node[0-9]*")):
    print(n, open(n + "/cpulist").read().strip())
try:
    print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout)
except Exception as e:
    print("topo:", e)
bdf = torch.cuda.get_device_properties(0)
pci = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip().splitlines()[0]
pci = pci.lower().replace("00000000:", "0000:")
try:
    print("gpu", pci, "numa_node", open(f"/sys/bus/pci/devices/{pci}/numa_node").read().strip())
except Exception as e:
    print("numa_node:", e)

dev = torch.device("cuda", 0)
n = 8_640_000 // 8
dst = torch.empty(n, dtype=torch.float64, device=dev)
full = sorted(os.sched_getaffinity(0))


def bw(tag):
    h = torch.empty(n, dtype=torch.float64).pin_memory()
    h.fill_(1.0)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(5):
            dst.copy_(h, non_blocking=True)
        s.synchronize()
        t0 = time.perf_counter()
        for _ in range(50):
            dst.copy_(h, non_blocking=True)
        s.synchronize()
        dt = (time.perf_counter() - t0) / 50
    print(f"{tag}: {n * 8 / dt / 1e9:.1f} GB/s ({dt * 1e6:.0f} us per 8.64 MB)")


bw("default")
for nd in sorted(glob.glob("/sys/devices/system/node/node[0-9]*")):
    cl = open(nd + "/cpulist").read().strip()
    cpus = set()
    for part in cl.split(","):
        if "-" in part:
            a, b = part.split("-"); cpus |= set(range(int(a), int(b) + 1))
        elif part:
            cpus.add(int(part))
    ok = sorted(cpus & set(full))
    if not ok:
        print(nd, "no allowed cpus"); continue
    os.sched_setaffinity(0, ok)
    bw(os.path.basename(nd))
os.sched_setaffinity(0, full)
