#!/usr/bin/env python3
"""NUMA placement of pinned staging on a multi-GPU box: where the GPUs sit,
what this process may use, and the aggregate pinned-copy rate of all GPUs at
once with every buffer on its GPU's node (soda_cuda_host_alloc) against every
buffer allocated for GPU 0's node.  One JSON line per measurement."""
import ctypes
import glob
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402


def read(path):
  try:
    with open(path) as fp:
      return fp.read().strip()
  except OSError as e:
    return 'n/a (%s)' % e.__class__.__name__


def main():
  n = torch.cuda.device_count()
  info = {'gpus': n, 'cpus_allowed': len(os.sched_getaffinity(0)),
          'nodes': {}, 'gpu_node': {}}
  for node in sorted(glob.glob('/sys/devices/system/node/node[0-9]*')):
    info['nodes'][os.path.basename(node)] = {
        'cpulist': read(node + '/cpulist'),
        'mem': read(node + '/meminfo').split('\n')[0][-24:]}
  for line in read('/proc/self/status').split('\n'):
    if line.startswith(('Mems_allowed_list', 'Cpus_allowed_list')):
      info[line.split(':')[0]] = line.split(':')[1].strip()
  for g in range(n):
    bus = torch.cuda.get_device_properties(g).pci_bus_id if hasattr(
        torch.cuda.get_device_properties(g), 'pci_bus_id') else None
    info['gpu_node'][g] = bus
  print(json.dumps(info), flush=True)

  st, prog = bench.config_program(bench.HEADLINE)
  lib = prog.lib
  lib.soda_cuda_host_alloc.argtypes = [ctypes.POINTER(ctypes.c_void_p),
                                       ctypes.c_int64, ctypes.c_int32]
  lib.soda_cuda_host_free.argtypes = [ctypes.c_void_p, ctypes.c_int64]
  rt = ctypes.CDLL('libcudart.so.12')
  rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p,
                                 ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
  nbytes = 512 << 20

  def measure(alloc_device_of):
    hosts, devs, streams = [], [], []
    for g in range(n):
      p_in, p_out = ctypes.c_void_p(), ctypes.c_void_p()
      assert lib.soda_cuda_host_alloc(ctypes.byref(p_in), nbytes,
                                      alloc_device_of(g)) == 0
      assert lib.soda_cuda_host_alloc(ctypes.byref(p_out), nbytes,
                                      alloc_device_of(g)) == 0
      hosts.append((p_in, p_out))
      with torch.cuda.device(g):
        devs.append((torch.empty(nbytes, dtype=torch.uint8, device='cuda'),
                     torch.ones(nbytes, dtype=torch.uint8, device='cuda')))
        streams.append((torch.cuda.Stream(), torch.cuda.Stream()))
    out = {}
    for mode in ('h2d', 'd2h', 'both'):
      best = None
      for rep in range(4):
        for g in range(n):
          torch.cuda.synchronize(g)
        t0 = time.perf_counter()
        for g in range(n):
          with torch.cuda.device(g):
            if mode in ('h2d', 'both'):
              rt.cudaMemcpyAsync(devs[g][0].data_ptr(), hosts[g][0], nbytes, 1,
                                 streams[g][0].cuda_stream)
            if mode in ('d2h', 'both'):
              rt.cudaMemcpyAsync(hosts[g][1], devs[g][1].data_ptr(), nbytes, 2,
                                 streams[g][1].cuda_stream)
        for g in range(n):
          torch.cuda.synchronize(g)
        dt = time.perf_counter() - t0
        if rep:
          best = dt if best is None else min(best, dt)
      out[mode] = round(n * nbytes * (2 if mode == 'both' else 1) / best / 1e9, 1)
    for p_in, p_out in hosts:
      lib.soda_cuda_host_free(p_in, nbytes)
      lib.soda_cuda_host_free(p_out, nbytes)
    return out

  print(json.dumps({'placement': 'own node (soda_cuda_host_alloc per GPU)',
                    'aggregate_gbs': measure(lambda g: g)}), flush=True)
  print(json.dumps({'placement': 'all on the node of GPU 0',
                    'aggregate_gbs': measure(lambda g: 0)}), flush=True)
  print(json.dumps({'placement': 'all on the node of the last GPU',
                    'aggregate_gbs': measure(lambda g: n - 1)}), flush=True)
  print(json.dumps({'placement': 'no preference (device -1)',
                    'aggregate_gbs': measure(lambda g: -1)}), flush=True)


if __name__ == '__main__':
  main()
