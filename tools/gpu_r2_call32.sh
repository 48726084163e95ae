#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
for lib in "" gpurun_variants/libevw_gnmb3.so gpurun_variants/libevw_gnmb4.so; do
  echo "== ${lib:-default}"
  if [ -n "$lib" ]; then export EVW_LIB=$PWD/$lib; else unset EVW_LIB; fi
  timeout 300 python - <<'PY'
import torch, sys
sys.path.insert(0, ".")
from evoworld_b200 import ops
dev = torch.device("cuda:0")
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for C, rows in [(320, 28 * 9216), (640, 28 * 9216), (640, 28 * 2304), (1280, 28 * 2304)]:
    x = torch.randn(rows, C, device=dev); g = torch.ones(C, device=dev); b = torch.zeros(C, device=dev)
    ms = timeit(lambda: ops.group_norm(x, g, b, 28, 1e-6, True))
    print(f"GN(28) stats+apply rows={rows} C={C}: {ms*1e3:7.1f} us  {rows*C*10/ms/1e6:7.0f} GB/s")
PY
done
