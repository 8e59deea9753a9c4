"""Probe what the box exposes about GPU <-> CPU / NUMA locality (the sysfs numa_node of the GPUs is -1 on this pool)."""
import os
import subprocess

print(subprocess.run(['nvidia-smi', 'topo', '-m'], capture_output=True, text=True).stdout)
print(subprocess.run('lscpu | grep -i -E "numa|socket|model name|^CPU\\(s\\)"', shell=True, capture_output=True, text=True).stdout)
print('nodes:', os.listdir('/sys/devices/system/node') if os.path.isdir('/sys/devices/system/node') else None)
for n in sorted(os.listdir('/sys/devices/system/node')):
    if n.startswith('node'):
        print(n, open(f'/sys/devices/system/node/{n}/cpulist').read().strip(), open(f'/sys/devices/system/node/{n}/meminfo').read().split('\n')[0])
print('affinity of this process:', len(os.sched_getaffinity(0)), sorted(os.sched_getaffinity(0))[:8], '...')
try:
    import pynvml
    pynvml.nvmlInit()
    for i in range(pynvml.nvmlDeviceGetCount()):
        h = pynvml.nvmlDeviceGetHandleByIndex(i)
        pci = pynvml.nvmlDeviceGetPciInfo(h).busId
        try:
            aff = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
            aff = [hex(int(a)) for a in aff]
        except Exception as e:  # noqa: BLE001
            aff = repr(e)
        try:
            mem = pynvml.nvmlDeviceGetMemoryAffinity(h, 4, 0)
            mem = [hex(int(a)) for a in mem]
        except Exception as e:  # noqa: BLE001
            mem = repr(e)
        node = None
        p = '/sys/bus/pci/devices/%s/numa_node' % (pci.decode() if isinstance(pci, bytes) else pci).lower()[-12:]
        if os.path.exists(p):
            node = open(p).read().strip()
        print('gpu', i, pci, 'cpu affinity', aff, 'mem affinity', mem, 'sysfs numa_node', node)
except Exception as e:  # noqa: BLE001
    print('pynvml:', repr(e))
