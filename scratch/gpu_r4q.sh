#!/bin/bash
# NUMA placement of the host threads / pinned buffers: first bench process on a fresh box
set -x
mkdir -p gpurun_out
python - <<'PY'
import torch, os, glob
p = torch.cuda.get_device_properties(0)
bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
print("gpu", bdf, "numa_node", open("/sys/bus/pci/devices/%s/numa_node" % bdf).read().strip(), "allowed cpus", len(os.sched_getaffinity(0)),
      "nodes", [(os.path.basename(d), open(d + "/cpulist").read().strip()) for d in sorted(glob.glob("/sys/devices/system/node/node[0-9]*"))])
PY
timeout 300 python bench.py --no-cpu --no-others --steps 3 --warmup 3 > gpurun_out/r4q_bench.json 2> gpurun_out/r4q_bench.err; tail -n 2 gpurun_out/r4q_bench.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r4q_bench.json") if l.startswith("{")][0])
print(d["value"], d["e2e"]["value"], d["config"].get("host_affinity"), "fwd", d["forward"]["value"], d["forward"]["e2e"], {k: v["value"] for k, v in d["forward"]["e2e_wire"].items()})
PY
