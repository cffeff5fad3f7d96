"""Training path (BASELINE config 4 building block): forward with batch-statistics BatchNorm and the
backward of every parameter, against PyTorch autograd through the same layers on the CPU
(oracle/torch_port.forward_train - the ATen ops the reference module issues in .train() mode)."""
import numpy as np
import pytest
import torch

from oracle import torch_port
from voicesplit_b200 import config, synth

pytestmark = pytest.mark.gpu


def _cpu_reference(sd_np, x, emb, gw, activation):
    sd = {}
    for k, v in sd_np.items():
        t = torch.from_numpy(np.array(v))
        if t.dtype == torch.float32 and "running" not in k:
            t.requires_grad_(True)
        sd[k] = t
    et = torch.from_numpy(emb).requires_grad_(True)
    mask = torch_port.forward_train(sd, torch.from_numpy(x), et, activation)
    (mask * torch.from_numpy(gw)).sum().backward()
    return mask.detach(), sd, et.grad


@pytest.mark.parametrize("tensor_cores", [True, False])
@pytest.mark.parametrize("model_name,dims,B,T", [("voicesplit", synth.make_dims(33, 16, 24, 40), 3, 21),
                                                 ("voicefilter", synth.make_dims(17, 8, 16, 24), 2, 9),
                                                 ("voicesplit", synth.make_dims(41, 20, 28, 36), 2, 40)])
def test_train_forward_backward_match_autograd(model_name, dims, B, T, tensor_cores):
    from models.voicefilter.model import VoiceFilter
    from models.voicesplit.model import VoiceSplit
    sd_np = synth.make_state_dict(dims, 5, "stress")
    x, emb = synth.make_inputs(B, T, dims, 6)
    gw = np.random.default_rng(0).standard_normal((B, T, dims["num_freq"])).astype(np.float32)
    ref_mask, ref_sd, ref_gemb = _cpu_reference(sd_np, x, emb, gw, model_name)

    cls = VoiceSplit if model_name == "voicesplit" else VoiceFilter
    m = cls(config.AttrDict(synth.make_config_dict(dims, model_name)))
    m.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd_np.items()})
    m = m.cuda().train()
    m.train_tensor_cores = tensor_cores
    et = torch.from_numpy(emb).cuda().requires_grad_(True)
    mask = m(torch.from_numpy(x).cuda(), et)
    assert mask.requires_grad
    (mask * torch.from_numpy(gw).cuda()).sum().backward()
    torch.cuda.synchronize()

    assert (mask.detach().cpu() - ref_mask).abs().max() < 2e-4
    msd = m.state_dict()
    for k in msd:                                    # running statistics updated like nn.BatchNorm2d
        if "running" in k:
            assert torch.allclose(msd[k].cpu(), ref_sd[k], atol=2e-5, rtol=1e-4), k
        if "num_batches" in k:
            assert int(msd[k]) == 1
    # relative to each gradient's own scale; a conv bias in front of a BatchNorm has an analytically ZERO
    # gradient (the batch mean removes it), so both sides hold rounding noise there: absolute floor
    gmax = max(float(ref_sd[k].grad.abs().max()) for k, _ in m.named_parameters())
    worst = {}
    for k, p in m.named_parameters():
        assert p.grad is not None, k
        r = ref_sd[k].grad
        err = float((p.grad.cpu() - r).abs().max())
        worst[k] = (err, float(r.abs().max()))
    conv_bias = {f"conv.{i}.bias": f"conv.{i}.weight" for i in (1, 5, 9, 13, 17, 21, 25, 28)}

    def scale(k):   # analytically-zero conv-bias gradients are judged on the scale of their layer's weight gradient
        return max(worst[k][1], worst[conv_bias[k]][1]) if k in conv_bias else worst[k][1]
    bad = {k: v for k, v in worst.items() if v[0] > 2e-3 * scale(k) + 1e-6 * gmax}
    print({k: (f"{v[0]:.2e}", f"{v[1]:.2e}") for k, v in worst.items()})
    assert not bad, bad
    assert (et.grad.cpu() - ref_gemb).abs().max() < 2e-3 * max(1e-6, float(ref_gemb.abs().max()))


def test_adam_step_changes_the_mask_and_repacks():
    """train.py:109-111: zero_grad / backward / step on the module's own parameters."""
    from models.voicesplit.model import VoiceSplit
    dims = synth.make_dims(17, 8, 16, 24)
    m = VoiceSplit(config.AttrDict(synth.make_config_dict(dims))).cuda().train()
    opt = torch.optim.Adam(m.parameters(), lr=1e-2)
    x, emb = synth.make_inputs(2, 12, dims, 1)
    xt, et = torch.from_numpy(x).cuda(), torch.from_numpy(emb).cuda()
    target = torch.rand_like(xt)
    losses = []
    for _ in range(6):
        opt.zero_grad()
        mask = m(xt, et)
        loss = ((mask * xt - target * xt) ** 2).mean()
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert losses[-1] < losses[0]
    m.eval()
    with torch.no_grad():
        assert torch.isfinite(m(xt, et)).all()
