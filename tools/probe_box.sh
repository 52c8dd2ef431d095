mkdir -p gpurun_out
{
echo "== nvidia-smi topo"; nvidia-smi topo -m
echo "== nodes"; ls /sys/devices/system/node/; cat /sys/devices/system/node/online
echo "== lscpu"; lscpu | head -40
echo "== cpuset"; cat /proc/self/status | grep -i -E "cpus_allowed_list|mems_allowed_list"
echo "== meminfo"; head -5 /proc/meminfo
echo "== numactl"; which numactl && numactl -H
echo "== nproc"; nproc
} > gpurun_out/r02_topology.txt 2>&1
tools/bin/h2d_ceiling 8 1024 4 > gpurun_out/r02_h2d_ceiling.json 2> gpurun_out/r02_h2d_ceiling.err
cat gpurun_out/r02_h2d_ceiling.json
