# round 2, GPU call M (1 GPU): which launch attribute makes ncu refuse the resident kernel; single-pass pre-passes (stream_map)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in "SWB_NO_COOP=1" "SWB_NO_L2_PERSIST=1" "SWB_NO_COOP=1 SWB_NO_L2_PERSIST=1"; do
  ( env $v timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/smoke_ncu_m.csv python __graft_entry__.py --smoke > gpurun_out/smoke_ncu_m.log 2>&1; echo "smoke under ncu with $v: rc=$?"; tail -1 gpurun_out/smoke_ncu_m.log; grep -v '^"' gpurun_out/smoke_ncu_m.csv | grep -i "error" | head -3; grep "merge_cluster" gpurun_out/smoke_ncu_m.csv | rev | cut -d, -f1-3 | rev )
done
( timeout 900 python -m pytest tests/test_pretok.py tests/test_normalize.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -15 > gpurun_out/pytest_m.log; tail -6 gpurun_out/pytest_m.log )
timeout 600 python - <<'PY'
import sys, time, ctypes
sys.path.insert(0, ".")
import numpy as np, torch
from shredword_b200 import synth
from shredword_b200.cbase import lib
arr = synth.corpus_bytes(synth.small_spec(1_000_000_000, 2_000_000, 11))
d = torch.from_numpy(arr).cuda()
out = torch.empty(2 * d.numel() + 16, dtype=torch.uint8, device="cuda")
for name, fn in (("pretokenize", lib.swb_pretokenize), ("normalize", lib.swb_normalize)):
  for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    n = fn(ctypes.c_void_p(d.data_ptr()), d.numel(), ctypes.c_void_p(out.data_ptr()), out.numel(), 1)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
  print(f"{name}: {d.numel()} -> {n} bytes in {dt*1e3:.1f} ms = {d.numel()/1e9/dt:.1f} GB/s (device resident, 1 GB ascii)")
PY
