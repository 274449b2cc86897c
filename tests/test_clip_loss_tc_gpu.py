"""Tensor-core variant of the contrastive loss (CTK_CLIP_LOSS_TC=1, N >= 1024; head.cu `clip_loss_tc`, GEMM epilogues
LSE_PART / CLIP_GRAD) against the fp64 oracle and against the default fp32 SIMT path.

Validated on a B200 in round 2 (gpurun_out/r2_clip_loss_tc.log: 6 passed).
The switch is read once per process, so each case runs in a child process with the variable set.
Tolerances: split-bf16 operands carry 16 mantissa bits, so logits are exact to ~2^-15 * exp(log_temp): loss 1e-4
relative, latent gradients 1e-3 of their largest entry (the fp32 SIMT path holds 1e-5 / 1e-4).
"""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu

CHILD = r"""
import json, sys, torch
sys.path.insert(0, sys.argv[1])
from oracle import ctclip_oracle as orc
from vit_exp_b200 import ops
N, W, rank, lt_val = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), float(sys.argv[5])
d, B = 512, N // W
g = torch.Generator().manual_seed(0)
T = torch.nn.functional.normalize(torch.randn(N, d, generator=g), dim=-1)
I = torch.nn.functional.normalize(torch.randn(N, d, generator=g) + 0.5 * T, dim=-1)
lt = torch.tensor(lt_val)
ref = orc.clip_loss_and_local_grads(T.double(), I.double(), lt.double(), B, rank)
dev = torch.device("cuda:0")
out, dl = ops.clip_loss_fwd_bwd(T.to(dev), I.to(dev), lt.reshape(1).to(dev), b_local=B, row0=rank * B)
torch.cuda.synchronize()
out, dl = out.cpu().double(), dl.cpu().double()
res = {"loss_rel": float(abs(out[0] - ref["loss"]) / abs(ref["loss"])),
       "dtemp_rel": float(abs(out[1] - ref["dlog_temp"]) / (abs(ref["dlog_temp"]) + 1e-12)),
       "dT_rel": float((dl[0] - ref["dT_local"]).abs().max() / ref["dT_local"].abs().max()),
       "dI_rel": float((dl[1] - ref["dI_local"]).abs().max() / ref["dI_local"].abs().max()),
       "finite": bool(torch.isfinite(out).all() and torch.isfinite(dl).all())}
print("RESULT " + json.dumps(res))
"""


def _run(N, W, rank, lt, tc):
    env = dict(os.environ, CTK_CLIP_LOSS_TC="1" if tc else "0")
    r = subprocess.run([sys.executable, "-c", CHILD, ROOT, str(N), str(W), str(rank), str(lt)], env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1]
    return json.loads(line[len("RESULT "):])


@pytest.mark.parametrize("N,W,rank,lt", [(1024, 1, 0, 1.0), (2048, 8, 3, 1.0), (4096, 8, 7, 1.0), (4096, 1, 0, 2.5),
                                         (1152, 4, 1, 0.0)])
def test_tc_loss_vs_oracle(N, W, rank, lt):
    res = _run(N, W, rank, lt, tc=True)
    assert res["finite"]
    assert res["loss_rel"] < 1e-4, res
    assert res["dtemp_rel"] < 1e-3, res
    assert res["dT_rel"] < 1e-3 and res["dI_rel"] < 1e-3, res


def test_simt_path_unchanged_when_switch_is_off():
    res = _run(1024, 4, 2, 1.0, tc=False)
    assert res["loss_rel"] < 1e-5 and res["dT_rel"] < 1e-4 and res["dI_rel"] < 1e-4, res
