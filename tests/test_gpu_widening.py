"""SURVEY §8(f) on the GPU: the opt-in batched `Modified_UNET` entry (row 1), checkpoint interchange + the tlit-style
no-grad evaluation loop (row 4).  Everything runs through the public module API (the custom ops over libgwn)."""
import os

import numpy as np
import pytest
import torch

from oracle.cases import GOLDEN_CASES, GOLDEN_DIR, case_inputs, case_supports, synthetic_module_state
from oracle.gwnet_oracle import synthetic_state_dict
from gpu_helpers import rel

pytestmark = pytest.mark.gpu


def _front_end(horizon):
    from multimodal_outage_b200.unet import UNetGWNetEncoder
    torch.manual_seed(3)
    m = UNetGWNetEncoder(horizon, input_channels=1, device='cuda').cuda()
    shapes = [(k, tuple(v.shape), str(v.dtype)) for k, v in m.state_dict().items() if not k.startswith('st_gnn.')]
    m.load_state_dict(synthetic_module_state(shapes, 5), strict=False)
    return m


def test_batched_gwnet_entry_matches_the_literal_per_sample_calls_when_batchnorm_is_frozen():
    """unet.py:221-225 calls gwnet once per sample with a [67,h,320] buffer.  `batched=True` stacks the B buffers into ONE
    [B,320,67,h] call (each sample reinterpreted exactly as the literal call does): eval mode -> same numbers; the
    literal path itself is bit-identical to calling gwnet directly on each sample's buffer (it is untouched)."""
    h, B = 3, 3
    m = _front_end(h).eval()
    g = torch.Generator(device='cpu').manual_seed(1)
    tiles = torch.randn(B, 67, h, 1, 128, 128, generator=g).cuda()
    tdim = torch.randn(B, 67, h, 64, generator=g).cuda()
    with torch.no_grad():
        lit = m(tiles, tdim)                                   # literal: per-sample gwnet calls
        bat = m(tiles, tdim, batched=True, literal_features=True)
        x = m.features(tiles, tdim, literal=True).contiguous()
        direct = torch.stack([m.st_gnn(x[b]) for b in range(B)])
    assert tuple(lit.shape) == (B, 67, h, 256)
    assert torch.equal(lit, direct)
    assert rel(bat, lit) < 1e-5, rel(bat, lit)
    # batched feature producer (counties and samples in one conv batch): same function in eval mode
    with torch.no_grad():
        bat2 = m(tiles, tdim, batched=True)
    assert rel(bat2, lit) < 1e-4, rel(bat2, lit)
    # training mode: BatchNorm spans the batch in the batched call - results differ from the per-sample schedule,
    # which is why it is opt-in (and gradients flow through the stacked call)
    m.train()
    m.st_gnn.dropout = 0.0
    out_b = m(tiles[:2], tdim[:2], batched=True)
    out_l = m(tiles[:2], tdim[:2])
    assert rel(out_b, out_l) > 1e-3
    out_b.square().mean().backward()
    assert m.contraction.inc.double_conv[0].weight.grad is not None
    assert m.st_gnn.start_conv.weight.grad is not None


def test_reference_checkpoint_on_disk_then_tlit_style_eval_loop_fp32_and_bf16(tmp_path):
    """lit.py:187-196 / tlit.py:46-94: load a Lightning-format checkpoint whose gwnet lives under `model.st_gnn.`, then run
    the no-grad evaluation loop.  The weights are the ones the reference produced `out_eval` with (literal golden:
    the reference's own classes executed by make_golden.py) - fp32 <= 1e-4, bf16 (running statistics, dropout off,
    tensor-core kernels) <= 2e-2 against the reference's output."""
    from multimodal_outage_b200 import gwnet
    from multimodal_outage_b200.unet import load_reference_checkpoint
    c = GOLDEN_CASES['literal']
    cfg, g = c['cfg'], np.load(os.path.join(GOLDEN_DIR, 'literal.npz'))
    keys = [str(k) for k in g['state_keys']]
    sd = synthetic_state_dict(cfg, c['seed'])
    # the running statistics the reference had when it produced out_eval (after its one training step)
    for k in g.files:
        if k.startswith('buf/'):
            sd[k[4:]] = torch.tensor(g[k])
    path = os.path.join(tmp_path, 'epoch=3-step=12.ckpt')
    torch.save({'epoch': 3, 'global_step': 12, 'state_dict': {f'model.st_gnn.{k}': sd[k] for k in keys}}, path)
    m = gwnet('cuda', num_nodes=67, in_dim=320, out_dim=256, horizon=c['horizon'],
              supports=[torch.tensor(s) for s in case_supports(c['supports'])])
    res = load_reference_checkpoint(m, path, prefix='model.st_gnn.')
    assert not res.missing_keys and not res.unexpected_keys
    m.eval()
    x_np, _ = case_inputs('literal')
    x = torch.tensor(x_np, device='cuda')
    ref = torch.tensor(g['out_eval'])
    with torch.no_grad():
        out32 = m(x)
        assert tuple(out32.shape) == tuple(ref.shape)
        assert rel(out32, ref) < 1e-4, rel(out32, ref)
        # the evaluation loop: several "batches", bf16 autocast like a mixed-precision test script
        tracked = int(m.bn[0].num_batches_tracked)
        for _ in range(3):
            with torch.autocast('cuda', dtype=torch.bfloat16):
                out16 = m(x)
            assert out16.dtype == torch.float32
            assert rel(out16, ref) < 2e-2, rel(out16, ref)
        assert int(m.bn[0].num_batches_tracked) == tracked          # eval mode leaves the statistics alone
