for L in 32 16 8; do for N in 4096 64; do echo "lanes=$L n=$N"; CKZG_B200_HASH_LANES=$L PROBE_N=$N timeout 120 python tools/gpu_probe.py modes 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.readline())
print(' dev',d['device_concurrent_ms'],'host',d['host_concurrent_ms'], {k:v for k,v in d['device_level1_stages_ms'].items() if 't_' in k or 'per_blob' in k}, {k:v for k,v in d['device_serial_level2_kernels_ms'].items() if k in ('hash+validate','evaluate')})
"; done; done
