"""GPU suite, last file on purpose: the persistent GEMM next to other work on the same GPU.

fc_gemm_kernel spins on device-side flags (split-K owners wait for their peers) and on a grid-wide barrier (finish
phase), which is only safe if all of its CTAs are co-resident.  It is therefore launched cooperatively
(azn_launch_coop, csrc/common.cuh): two such grids issued on two streams at the same time must be placed one after the
other by the driver instead of interleaving halves of each and deadlocking.  The scenario runs in a child process under
a timeout so that a regression shows up as a failed test, not as a hung box."""
import subprocess
import sys
import os

import pytest

torch = pytest.importorskip("torch")

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import sys
sys.path.insert(0, %r)
import torch
from aznet_b200 import _lib as L, ops
L.build(); L.require_device()
dev = torch.device("cuda:0")
gen = torch.Generator(device="cpu").manual_seed(3)
K, N = 25088, 4096
W = (torch.randn((N, K), generator=gen) * 0.01).to(torch.bfloat16).to(dev)
b = torch.zeros(N, device=dev)
cases = []
for M in (64, 600):                       # M = 64: split-K with the grid-barrier finish phase; 600: owner/peer flags
    A = torch.relu(torch.randn((M, K), generator=gen)).to(torch.bfloat16).to(dev)
    ref = ops.fc_forward(A, W, b, L.ACT_RELU).clone()
    cases.append((A, ref))
torch.cuda.synchronize()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
outs = []
for it in range(12):
    for k, st in enumerate((s1, s2)):
        A, ref = cases[(it + k) %% 2]
        with torch.cuda.stream(st):
            outs.append((ops.fc_forward(A, W, b, L.ACT_RELU), ref))
    if it %% 3 == 0:                      # a long elementwise kernel in between, on the default stream
        torch.empty(64 << 20, device=dev).normal_()
torch.cuda.synchronize()
assert all(torch.equal(o, r) for o, r in outs), "concurrent launches changed a result"
print("CONCURRENT-OK")
'''


def test_two_persistent_gemms_on_two_streams_do_not_deadlock():
    from aznet_b200 import _lib
    _lib.build()
    _lib.require_device()
    try:
        p = subprocess.run([sys.executable, "-c", CHILD % ROOT], capture_output=True, text=True, timeout=180)
    except subprocess.TimeoutExpired:
        pytest.fail("two concurrent azn_fc_forward launches did not finish within 180 s (deadlock)")
    assert p.returncode == 0 and "CONCURRENT-OK" in p.stdout, (p.stdout[-2000:], p.stderr[-2000:])
