import os, time, subprocess, torch
print(subprocess.run("nvidia-smi topo -m | head -8; lscpu | grep -i -E 'numa|^CPU\\(s\\)|Model name'; nproc", shell=True, capture_output=True, text=True).stdout)
print("affinity", sorted(os.sched_getaffinity(0))[:40], len(os.sched_getaffinity(0)))
dev = torch.device("cuda:0")
n = 166 << 20
def bw(tag):
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    ts = []
    for i in range(6):
        t0 = time.perf_counter(); d.copy_(h, non_blocking=True); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    print(tag, "H2D GB/s:", [round(n / t / 1e9, 1) for t in ts])
    h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
    t0 = time.perf_counter(); h2.copy_(d, non_blocking=True); torch.cuda.synchronize(); print(tag, "D2H GB/s", round(n / (time.perf_counter() - t0) / 1e9, 1))
bw("default")
cpus = sorted(os.sched_getaffinity(0))
for name, sel in (("first-half", cpus[: len(cpus) // 2]), ("second-half", cpus[len(cpus) // 2:])):
    os.sched_setaffinity(0, sel)
    bw(name)
