"""SURVEY §8(f) rows 1, 3, 4 on CPU: the batched UNet encoder front end against vectors produced by executing the
reference's own classes (tests/golden/make_golden_unet.py), and the on-disk checkpoint interchange."""
import os

import numpy as np
import pytest
import torch

from oracle.cases import GOLDEN_DIR, synthetic_module_state

G = np.load(os.path.join(GOLDEN_DIR, 'unet_enc.npz'))
N_COUNTIES, HORIZON, CHANNELS, SEED = 3, 2, 1, 77


def _shapes(m):
    return [(k, tuple(v.shape), str(v.dtype)) for k, v in m.state_dict().items()]


def _modules():
    from multimodal_outage_b200.unet import Contraction, Encoder
    con, enc = Contraction(CHANNELS, HORIZON), Encoder()
    con.load_state_dict(synthetic_module_state(_shapes(con), SEED), strict=True)
    enc.load_state_dict(synthetic_module_state(_shapes(enc), SEED + 1), strict=True)
    rng = np.random.default_rng(SEED + 2)
    x = torch.tensor(rng.standard_normal((N_COUNTIES, HORIZON, CHANNELS, 128, 128)).astype(np.float32))
    return con, enc, x


def test_state_dict_keys_and_shapes_match_the_reference_classes():
    con, enc, _ = _modules()
    assert [k for k, _, _ in _shapes(con)] == list(G['con_keys'])
    assert [str(s) for _, s, _ in _shapes(con)] == list(G['con_shapes'])
    assert [k for k, _, _ in _shapes(enc)] == list(G['enc_keys'])
    assert [str(s) for _, s, _ in _shapes(enc)] == list(G['enc_shapes'])


@pytest.mark.parametrize('literal', [True, False])
def test_contraction_and_encoder_eval_match_reference(literal):
    """Eval mode: the county loop and the single batched conv stack are the same function."""
    con, enc, x = _modules()
    con.eval(); enc.eval()
    with torch.no_grad():
        y = con(x, literal=literal)
        f = enc(y, literal=literal)
    assert tuple(y.shape) == G['con_eval'].shape
    assert np.abs(y.numpy() - G['con_eval']).max() <= 2e-5 * np.abs(G['con_eval']).max()
    assert np.abs(f.numpy() - G['enc_eval']).max() <= 2e-5 * max(np.abs(G['enc_eval']).max(), 1e-6)
    assert [str(tuple(m.shape)) for m in con.feature_maps] == list(G['fmap_shapes'])
    sums = np.array([float(m.double().sum()) for m in con.feature_maps])
    assert np.allclose(sums, G['fmap_eval_sums'], rtol=1e-4)


def test_literal_training_schedule_matches_reference_and_batched_differs():
    """Training mode: literal = the reference's per-county BatchNorm batches (running statistics updated once per
    county); the batched form normalises over all counties at once - a documented semantic difference."""
    con, _, x = _modules()
    con.train()
    y = con(x, literal=True)
    assert np.abs(y.detach().numpy() - G['con_train']).max() <= 5e-5 * np.abs(G['con_train']).max()
    bn = con.inc.double_conv[1]
    assert int(bn.num_batches_tracked) == int(G['bn_tracked_after']) == N_COUNTIES
    assert np.allclose(bn.running_mean.numpy(), G['bn_running_mean_after'], atol=1e-6)
    con2, _, _ = _modules()
    con2.train()
    yb = con2(x)
    assert int(con2.inc.double_conv[1].num_batches_tracked) == 1
    assert (yb - y).abs().max() > 1e-3          # different normalisation batches


def test_batch_dimension_folds_into_the_same_call():
    con, enc, x = _modules()
    con.eval(); enc.eval()
    xb = torch.stack([x, x.flip(0)])
    with torch.no_grad():
        yb = enc(con(xb))
        assert tuple(con.feature_maps[0].shape[:3]) == (2, N_COUNTIES, HORIZON)
        y0, y1 = enc(con(xb[0])), enc(con(xb[1]))
    assert tuple(yb.shape) == (2, N_COUNTIES, HORIZON, 256)
    assert torch.allclose(yb[0], y0, atol=1e-5) and torch.allclose(yb[1], y1, atol=1e-5)


def test_reference_checkpoint_formats_load(tmp_path):
    """lit.py:187-196 writes Lightning checkpoints of LitModified_UNET: {'state_dict': {'model.<path>': tensor}}.
    A bare gwnet picks its entries with prefix 'model.st_gnn.'; keys of the decoder half are ignored; the module's own
    keys load strictly (key names and shapes are pinned to the reference by the golden `state_keys`)."""
    from multimodal_outage_b200 import gwnet
    from multimodal_outage_b200.unet import load_reference_checkpoint
    from oracle.gwnet_oracle import GWNetConfig, synthetic_state_dict
    cfg = GWNetConfig(num_nodes=67, in_dim=320, out_dim=256, kernel_size=1, n_fixed_supports=1, dropout=0.0)
    sd = synthetic_state_dict(cfg, 9)
    ref_keys = [str(k) for k in np.load(os.path.join(GOLDEN_DIR, 'literal.npz'))['state_keys']]
    assert sorted(k for k in sd if k in ref_keys) == sorted(ref_keys)          # the reference's own key set
    ckpt = {'epoch': 3, 'state_dict': {**{f'model.st_gnn.{k}': sd[k] for k in ref_keys},
                                       'model.decoder.fc1.weight': torch.zeros(4, 4),
                                       'model.contraction.inc.double_conv.0.weight': torch.zeros(4, 1, 3, 3)}}
    path = os.path.join(tmp_path, 'lit.ckpt')
    torch.save(ckpt, path)
    m = gwnet('cpu', in_dim=320, out_dim=256, horizon=3)
    res = load_reference_checkpoint(m, path, prefix='model.st_gnn.')
    assert not res.missing_keys and not res.unexpected_keys
    for k in ref_keys:
        assert torch.equal(m.state_dict()[k], sd[k]), k
    # plain state_dict file, no prefix
    torch.save({k: sd[k] for k in ref_keys}, path)
    m2 = gwnet('cpu', in_dim=320, out_dim=256, horizon=3)
    load_reference_checkpoint(m2, path, prefix='')
    assert torch.equal(m2.end_conv_2.weight, sd['end_conv_2.weight'])
    # a checkpoint that lacks one of the module's keys fails loudly (strict)
    bad = {k: v for k, v in sd.items() if k in ref_keys and k != 'start_conv.bias'}
    torch.save(bad, path)
    with pytest.raises(RuntimeError):
        load_reference_checkpoint(gwnet('cpu', in_dim=320, out_dim=256, horizon=3), path, prefix='')
