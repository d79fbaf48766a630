#!/bin/bash
# host-side bandwidth probe on the GPU-local and the remote NUMA node (profiles/probes/probe_host_expand.cu)
cd /root/repo
o=gpurun_out/r2a; mkdir -p $o
bus=$(nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader | head -1 | tr A-Z a-z | sed 's/^0000//')
node=$(cat /sys/bus/pci/devices/$bus/numa_node); echo "GPU0 bus $bus numa node $node"
local_cpus=$(cat /sys/devices/system/node/node$node/cpulist); other=$((1-node)); remote_cpus=$(cat /sys/devices/system/node/node$other/cpulist)
echo "== bound to GPU-local node $node ($local_cpus)"; timeout 300 taskset -c $local_cpus profiles/probes/bin/probe_host_expand > $o/probe_host_expand_local.txt 2>&1; tail -22 $o/probe_host_expand_local.txt
echo "== bound to remote node $other ($remote_cpus)"; timeout 300 taskset -c $remote_cpus profiles/probes/bin/probe_host_expand > $o/probe_host_expand_remote.txt 2>&1; tail -22 $o/probe_host_expand_remote.txt
echo "== unbound"; timeout 300 profiles/probes/bin/probe_host_expand > $o/probe_host_expand_unbound.txt 2>&1; tail -22 $o/probe_host_expand_unbound.txt
