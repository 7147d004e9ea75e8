#!/bin/bash
# round-2 GPU call d: the bound binary -- integration parity tests, whole-program timings on real Bifrost graphs, PCIe rates
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_integration.py tests/test_gpu_dropin.py -x -q > gpurun_out/r02d_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/r02d_rc.txt
python - > gpurun_out/r02d_pcie.txt 2>&1 <<'PY'
import torch, time
dev = torch.device("cuda:0")
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device=dev)
h2 = torch.empty(n, dtype=torch.uint8).pin_memory(); d2 = torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(f, reps=5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
print("H2D 1 GiB pinned: %.1f GB/s" % (n / t(lambda: d.copy_(h, non_blocking=True)) / 1e9))
print("D2H 1 GiB pinned: %.1f GB/s" % (n / t(lambda: h.copy_(d, non_blocking=True)) / 1e9))
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
print("H2D + D2H concurrently, each direction: %.1f GB/s" % (n / t(both) / 1e9))
p = torch.empty(n, dtype=torch.uint8)
print("H2D 1 GiB pageable: %.1f GB/s" % (n / t(lambda: d.copy_(p)) / 1e9))
PY
PF_PROGRAM_CHECK=1 timeout 1500 python integration/time_program.py 20000000 2 gpurun_out/r02d_prog_20m_dip.json > gpurun_out/r02d_prog_20m_dip.log 2>&1; echo "prog20 rc=$?" >> gpurun_out/r02d_rc.txt
timeout 2400 python integration/time_program.py 100000000 2 gpurun_out/r02d_prog_100m_dip.json > gpurun_out/r02d_prog_100m_dip.log 2>&1; echo "prog100 rc=$?" >> gpurun_out/r02d_rc.txt
tail -5 gpurun_out/r02d_tests.log; cat gpurun_out/r02d_rc.txt gpurun_out/r02d_pcie.txt; tail -c 3000 gpurun_out/r02d_prog_20m_dip.log; tail -c 3000 gpurun_out/r02d_prog_100m_dip.log
