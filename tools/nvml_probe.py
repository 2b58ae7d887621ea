import sys, time, threading, json
sys.path.insert(0, ".")
import torch, pynvml as n
import bench
n.nvmlInit(); h = n.nvmlDeviceGetHandleByIndex(0)
stop = False
res = {"sm": [], "max": [], "power": [], "reasons": []}
def sampler():
    while not stop:
        for name, fn in (("sm", lambda: n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)), ("max", lambda: n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM)),
                         ("power", lambda: n.nvmlDeviceGetPowerUsage(h)), ("reasons", lambda: n.nvmlDeviceGetCurrentClocksEventReasons(h))):
            t = time.perf_counter(); fn(); res[name].append(1e3 * (time.perf_counter() - t))
        time.sleep(0.2)
# GPU busy loop: matmuls with syncs, measuring sync jitter
x = torch.randn(8192, 8192, device="cuda", dtype=torch.float64)
th = threading.Thread(target=sampler); th.start()
lat = []
for i in range(40):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    y = x @ x
    e1.record(); torch.cuda.synchronize()
    lat.append((1e3 * (time.perf_counter() - t0), e0.elapsed_time(e1)))
stop = True; th.join()
print({k: (round(max(v), 2), round(sum(v) / len(v), 2), len(v)) for k, v in res.items()})
print("host-dev ms (max over steps):", max(a - b for a, b in lat), "median dev", sorted(b for a, b in lat)[20])
