"""TEST INFRASTRUCTURE - functional PyTorch (CPU) restatement of the reference forward pass.

The reference module is nothing but stock torch.nn layers
(/root/reference/models/voicesplit/model.py:15-64,66-89), so on a CPU its arithmetic is executed by
PyTorch's oneDNN/MKL kernels.  This file issues the same ATen ops from a plain state dict
(F.pad/conv2d/batch_norm, torch's LSTM, F.linear), which makes it (a) a second, independent
checker next to the C oracle and (b) the honest multi-threaded CPU baseline for bench.py
(`cpu_baseline`, `--impl reference`): the reference tree itself cannot travel to the GPU box.
Pinned by tests/test_oracle.py against the golden vectors of the unmodified reference.
Only tests/, smoke() and bench.py's baseline legs (CPU; plus the stock-PyTorch-on-the-same-GPU comparison of
SURVEY.md 8(d), where the same ATen ops dispatch to cuDNN/cuBLAS) may import it.
"""
import torch
import torch.nn.functional as F

_CONVS = ((1, 2, (3, 3, 0, 0), 1), (5, 6, (0, 0, 3, 3), 1), (9, 10, (2, 2, 2, 2), 1), (13, 14, (2, 2, 4, 4), 2),
          (17, 18, (2, 2, 8, 8), 4), (21, 22, (2, 2, 16, 16), 8), (25, 26, (2, 2, 32, 32), 16), (28, 29, None, 1))


def _act(x, kind):
    if kind in ("relu", "voicefilter"):
        return torch.relu(x)
    return x * torch.tanh(F.softplus(x))          # Mish, utils/generic_utils.py:399


def _t(sd, k):
    v = sd[k]
    return v if isinstance(v, torch.Tensor) else torch.from_numpy(v)


@torch.no_grad()
def conv_stack(sd, x, activation="mish"):
    """x [B,T,F] -> [B,T,8F] (index c*F+f)."""
    h = x.unsqueeze(1)
    for ci, bi, pad, dil in _CONVS:
        if pad is not None:
            h = F.pad(h, pad)
        h = F.conv2d(h, _t(sd, f"conv.{ci}.weight"), _t(sd, f"conv.{ci}.bias"), dilation=(dil, 1))
        h = F.batch_norm(h, _t(sd, f"conv.{bi}.running_mean"), _t(sd, f"conv.{bi}.running_var"),
                         _t(sd, f"conv.{bi}.weight"), _t(sd, f"conv.{bi}.bias"), training=False, eps=1e-5)
        h = _act(h, activation)
    B, C, T, Fq = h.shape
    return h.permute(0, 2, 1, 3).reshape(B, T, C * Fq)


@torch.no_grad()
def forward(sd, x, emb, activation="mish"):
    """x [B,T,F], emb [B,E] (torch CPU tensors or numpy) -> mask [B,T,F]."""
    x = x if isinstance(x, torch.Tensor) else torch.from_numpy(x)
    emb = emb if isinstance(emb, torch.Tensor) else torch.from_numpy(emb)
    B, T, _ = x.shape
    feat = torch.cat((conv_stack(sd, x, activation), emb[:, None, :].expand(B, T, emb.shape[1])), dim=2)
    H = _t(sd, "lstm.weight_hh_l0").shape[1]
    flat = [_t(sd, f"lstm.{n}_l0{s}") for s in ("", "_reverse") for n in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]
    zeros = torch.zeros(2, B, H, device=feat.device)
    out, _, _ = torch._VF.lstm(feat, (zeros, zeros), flat, True, 1, 0.0, False, True, True)
    y = F.linear(torch.relu(out), _t(sd, "fc1.weight"), _t(sd, "fc1.bias"))
    y = F.linear(torch.relu(y), _t(sd, "fc2.weight"), _t(sd, "fc2.bias"))
    return torch.sigmoid(y)


def forward_train(sd, x, emb, activation="mish", momentum=0.1):
    """Training-mode forward with autograd (BatchNorm batch statistics, running-stat update in place on
    the tensors of `sd`), the way train.py:84,94 drives the reference module.  `sd` maps the reference
    state_dict keys to torch tensors; parameters that should receive .grad must have requires_grad=True."""
    B, T, _ = x.shape
    h = x.unsqueeze(1)
    for ci, bi, pad, dil in _CONVS:
        if pad is not None:
            h = F.pad(h, pad)
        h = F.conv2d(h, sd[f"conv.{ci}.weight"], sd[f"conv.{ci}.bias"], dilation=(dil, 1))
        h = F.batch_norm(h, sd[f"conv.{bi}.running_mean"], sd[f"conv.{bi}.running_var"], sd[f"conv.{bi}.weight"],
                         sd[f"conv.{bi}.bias"], training=True, momentum=momentum, eps=1e-5)
        sd[f"conv.{bi}.num_batches_tracked"] += 1
        h = _act(h, activation)
    Bq, C, Tq, Fq = h.shape
    feat = torch.cat((h.permute(0, 2, 1, 3).reshape(B, T, C * Fq), emb[:, None, :].expand(B, T, emb.shape[1])), dim=2)
    H = sd["lstm.weight_hh_l0"].shape[1]
    flat = [sd[f"lstm.{n}_l0{s}"] for s in ("", "_reverse") for n in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]
    zeros = torch.zeros(2, B, H)
    out, _, _ = torch._VF.lstm(feat, (zeros, zeros), flat, True, 1, 0.0, True, True, True)
    y = F.linear(torch.relu(out), sd["fc1.weight"], sd["fc1.bias"])
    y = F.linear(torch.relu(y), sd["fc2.weight"], sd["fc2.bias"])
    return torch.sigmoid(y)
