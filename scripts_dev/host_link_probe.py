"""Host-link diagnosis: NUMA layout of the box, and H2D bandwidth of the two halves of one 4 GB pinned buffer."""
import glob, os, time
import torch
print("affinity:", sorted(os.sched_getaffinity(0)))
for n in sorted(glob.glob("/sys/devices/system/node/node*")):
    try:
        cl = open(n + "/cpulist").read().strip()
        mi = [l for l in open(n + "/meminfo") if "MemTotal" in l or "MemFree" in l]
        print(os.path.basename(n), "cpus", cl, "|", " ".join(" ".join(l.split()[2:]) for l in mi))
    except Exception as e:
        print(n, e)
bus = torch.cuda.get_device_properties(0).pci_bus_id if hasattr(torch.cuda.get_device_properties(0), "pci_bus_id") else None
for d in glob.glob("/sys/bus/pci/devices/*/numa_node"):
    try:
        cls = open(os.path.dirname(d) + "/class").read().strip()
        if cls.startswith("0x0302") or cls.startswith("0x0300"):
            print("gpu", os.path.dirname(d).split("/")[-1], "numa_node", open(d).read().strip())
    except Exception:
        pass
n = 1 << 30   # floats: 4 GB
h = torch.empty(n, dtype=torch.float32, pin_memory=True)
h.fill_(1.0)
dv = torch.empty(n // 2, dtype=torch.float32, device="cuda")
for rep in range(2):
    for name, sl in (("first half", slice(0, n // 2)), ("second half", slice(n // 2, n))):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        dv.copy_(h[sl], non_blocking=True)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print("pinned 4 GB buffer, %s: %.1f GB/s" % (name, 2.147 / dt))
# 16 x 4 MB chunks like the upload path
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    for name, base in (("first half chunks", 0), ("second half chunks", n // 2)):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for i in range(0, n // 2, 1 << 20):
            dv[i:i + (1 << 20)].copy_(h[base + i: base + i + (1 << 20)], non_blocking=True)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print("4 MB copies, %s: %.1f GB/s" % (name, 2.147 / dt))
