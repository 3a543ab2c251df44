import os, glob, time, torch
print("cpus", os.cpu_count(), "affinity", sorted(os.sched_getaffinity(0))[:4], "...", len(os.sched_getaffinity(0)))
for n in sorted(glob.glob("/sys/devices/system/node/node*")):
    print(n.split("/")[-1], open(n + "/cpulist").read().strip())
p = torch.cuda.get_device_properties(0)
try:
    bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
    print("gpu", bdf, "numa_node", open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
except Exception as e:
    print("pci info failed", e)
def bw(cpus):
    os.sched_setaffinity(0, cpus)
    h = torch.empty(64 << 20, dtype=torch.uint8, pin_memory=True); h.fill_(1)
    d = torch.empty_like(h, device="cuda")
    torch.cuda.synchronize()
    for _ in range(2): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    up = 10 * 64 / 1024 / (time.perf_counter() - t0)
    t0 = time.perf_counter()
    for _ in range(10): h.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    dn = 10 * 64 / 1024 / (time.perf_counter() - t0)
    return round(up, 1), round(dn, 1)
allc = sorted(os.sched_getaffinity(0))
half = len(allc) // 2
print("H2D/D2H GB/s, first half cpus", bw(set(allc[:half])))
print("H2D/D2H GB/s, second half cpus", bw(set(allc[half:])))
print("H2D/D2H GB/s, all cpus", bw(set(allc)))
