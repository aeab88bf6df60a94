"""vgl::MultiGpuSimulator (vcfgl_b200/host/vgl_host.hpp): several device contexts behind one driver loop,
consecutive batches (contiguous ranges of the global site index) alternating between them, records and gVCF blocks delivered
in site order by the ordered host-side merge.  Because every draw is keyed by the global site index, the output must be the
same bytes for any device list -- checked here on the C++ example driver: one context against two contexts on one GPU
(always), and against two GPUs (when the box has them), with and without the gVCF block machine stitching across batches.
"""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "vcfgl_b200", "host", "example_driver")


def n_gpus():
    import torch
    return torch.cuda.device_count()


def run(devices, n_sites, batch, gvcf):
    if not os.path.exists(EXE):
        pytest.skip("example_driver not built")
    env = dict(os.environ, VGL_DEVICES=devices, VGL_BATCH=str(batch))
    if gvcf:
        env.update(VGL_GVCF_DPS="1,3", VGL_INVARIANT="1")
    r = subprocess.run([EXE, str(n_sites)], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


@pytest.mark.parametrize("gvcf", [False, True], ids=["records", "gvcf_blocks"])
@pytest.mark.parametrize("batch", [8, 5])
def test_two_contexts_give_the_one_context_output(gvcf, batch):
    one = run("0", 203, batch, gvcf)
    assert one.count("\n") > 50 and ("BLOCK" in one) == gvcf
    assert run("0,0", 203, batch, gvcf) == one
    assert run("0,0,0", 203, batch, gvcf) == one
    assert run("0", 203, 64, gvcf) == one          # and it does not depend on the batch size either


@pytest.mark.parametrize("gvcf", [False, True], ids=["records", "gvcf_blocks"])
def test_two_gpus_give_the_one_gpu_output(gvcf):
    if n_gpus() < 2:
        pytest.skip("needs two GPUs")
    one = run("0", 403, 16, gvcf)
    assert run("0,1", 403, 16, gvcf) == one
    assert run(",".join(str(i) for i in range(n_gpus())), 403, 16, gvcf) == one
