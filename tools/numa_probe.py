"""Where is the GPU relative to the host NUMA nodes, and what does pinned-memory placement do to H2D?"""
import glob, os, subprocess, torch
print("affinity at start:", len(os.sched_getaffinity(0)), "cpus", sorted(os.sched_getaffinity(0))[:4], "...")
for p in sorted(glob.glob("/sys/devices/system/node/node*/cpulist")):
    print(p.split("/")[-2], open(p).read().strip())
try:
    import pynvml as nv
    nv.nvmlInit(); h = nv.nvmlDeviceGetHandleByIndex(0)
    words = (os.cpu_count() + 63) // 64
    mask = nv.nvmlDeviceGetCpuAffinity(h, words)
    cpus = sorted(64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1)
    print("nvml cpu affinity of GPU0:", len(cpus), cpus[:4], "...", cpus[-4:])
    bus = nv.nvmlDeviceGetPciInfo(h).busId
    bus = bus.decode() if isinstance(bus, bytes) else bus
    for cand in (bus.lower(), bus.lower()[4:]):
        f = f"/sys/bus/pci/devices/{cand}/numa_node"
        if os.path.exists(f): print("sysfs numa_node:", open(f).read().strip())
except Exception as e:
    print("nvml:", e)
dev = torch.device("cuda:0")
allcpus = sorted(os.sched_getaffinity(0))
def h2d(nbytes=32 << 20, reps=20):
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory(); h.fill_(1)
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    for _ in range(3): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    return reps * nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9
for p in sorted(glob.glob("/sys/devices/system/node/node*/cpulist")):
    node = p.split("/")[-2]
    cpus = set()
    for part in open(p).read().strip().split(","):
        a, _, b = part.partition("-"); cpus |= set(range(int(a), int(b or a) + 1))
    cpus &= set(allcpus)
    if not cpus: print(node, "no allowed cpus"); continue
    os.sched_setaffinity(0, cpus)
    print(f"{node}: allocate+touch pinned memory from its cpus -> H2D {h2d():.1f} {h2d():.1f} {h2d():.1f} GB/s", flush=True)
os.sched_setaffinity(0, allcpus)
print(f"unbound: H2D {h2d():.1f} {h2d():.1f} {h2d():.1f} GB/s")
