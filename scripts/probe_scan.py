"""Scan candidate (LBO, SBO) descriptor encodings for the tcgen05 probe, one subprocess per candidate
(a faulting candidate must not take the others down).  Writes gpurun_out/probe_scan.txt."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r"""
import sys, ctypes, torch
sys.path.insert(0, %r)
from prifit_b200 import _lib
mode, lbo, sbo = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
g = torch.Generator().manual_seed(5)
A = torch.randn(128, 128, generator=g).cuda()
Bm = torch.randn(128, 128, generator=g).cuda()
ws = torch.empty(40960, dtype=torch.uint8, device='cuda')
D = torch.zeros(128, 128, device='cuda')
_lib.call('prifit_debug_tc_probe', ctypes.c_void_p(A.data_ptr()), ctypes.c_void_p(Bm.data_ptr()), mode, lbo, sbo,
          ctypes.c_void_p(D.data_ptr()), ctypes.c_void_p(ws.data_ptr()), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
At = A.cpu().half().double(); Bt = Bm.cpu().half().double()
ref = At @ (Bt.T if mode == 0 else Bt)
print('RESULT mode %%d lbo %%d sbo %%d max_err %%.3e ref_max %%.2f' %% (mode, lbo, sbo, float((D.cpu().double()-ref).abs().max()), float(ref.abs().max())))
""" % ROOT

CANDIDATES = [(0, 16, 1024), (1, 16384, 1024), (1, 1024, 16384), (1, 16384, 2048), (1, 2048, 16384), (1, 16384, 128)]

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "probe_scan.txt"), "w") as out:
    for mode, lbo, sbo in CANDIDATES:
        try:
            r = subprocess.run([sys.executable, "-c", CHILD, str(mode), str(lbo), str(sbo)], capture_output=True, text=True, timeout=120)
            lines = [l for l in r.stdout.splitlines() if l.startswith("RESULT")]
            msg = lines[0] if lines else "FAILED mode %d lbo %d sbo %d rc=%d %s" % (mode, lbo, sbo, r.returncode, (r.stderr or "")[-300:].replace("\n", " | "))
        except subprocess.TimeoutExpired:
            msg = "TIMEOUT mode %d lbo %d sbo %d" % (mode, lbo, sbo)
        print(msg)
        out.write(msg + "\n")
        out.flush()
