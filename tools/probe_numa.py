#!/usr/bin/env python
"""Development probe: what the box shows about NUMA (nodes, their CPUs and memory, this process' affinity and memory
policy, the NUMA node of every GPU) — input for placing the pinned host buffers of dfx_*_host next to the GPU."""
import glob
import os
import subprocess


def rd(p):
    try:
        return open(p).read().strip()
    except OSError as e:
        return f"<{e.strerror}>"


print("affinity:", sorted(os.sched_getaffinity(0)))
for n in sorted(glob.glob("/sys/devices/system/node/node[0-9]*")):
    mem = [line for line in rd(n + "/meminfo").splitlines() if "MemTotal" in line or "MemFree" in line]
    print(os.path.basename(n), "cpus", rd(n + "/cpulist"), "|", " ; ".join(" ".join(m.split()[2:]) for m in mem))
print("online nodes:", rd("/sys/devices/system/node/online"), "| has_cpu:", rd("/sys/devices/system/node/has_cpu"),
      "| has_memory:", rd("/sys/devices/system/node/has_memory"))
out = subprocess.run(["nvidia-smi", "--query-gpu=index,pci.bus_id", "--format=csv,noheader"], capture_output=True, text=True).stdout
for line in out.strip().splitlines():
    idx, bus = [t.strip() for t in line.split(",")]
    dev = "/sys/bus/pci/devices/" + bus.lower()[4:]
    print("gpu", idx, bus, "numa_node", rd(dev + "/numa_node"), "local_cpulist", rd(dev + "/local_cpulist"))
print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[-2500:])
print("Mems_allowed:", [line for line in rd("/proc/self/status").splitlines() if "Mems_allowed" in line or "Cpus_allowed_list" in line])
try:
    import ctypes
    libc = ctypes.CDLL(None, use_errno=True)
    mode = ctypes.c_int(-1)
    mask = (ctypes.c_ulong * 16)()
    rc = libc.syscall(239, ctypes.byref(mode), mask, 1024, 0, 0)      # get_mempolicy
    print("get_mempolicy rc", rc, "errno", ctypes.get_errno(), "mode", mode.value, "mask", hex(mask[0]))
    mask2 = (ctypes.c_ulong * 16)()
    mask2[0] = 1
    rc = libc.syscall(238, 2, mask2, 1024)                              # set_mempolicy(MPOL_BIND, node 0)
    print("set_mempolicy(MPOL_BIND, node0) rc", rc, "errno", ctypes.get_errno())
    rc = libc.syscall(238, 0, None, 0)                                  # back to MPOL_DEFAULT
    print("set_mempolicy(MPOL_DEFAULT) rc", rc, "errno", ctypes.get_errno())
except Exception as e:  # noqa: BLE001
    print("mempolicy probe failed:", e)
