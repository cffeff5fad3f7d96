"""Data-parallel training over NCCL, 2 ranks on 2 GPUs of one box:
the gradients after `allreduce_gradients` must equal the single-process gradient of the CONCATENATED batch when
SyncBN is on (the reference is single-process: its BatchNorm statistics span the whole batch, train.py:84-111), and the
mean of the per-shard single-process gradients when every rank keeps its own statistics (the default)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from voicesplit_b200 import config, synth

pytestmark = pytest.mark.gpu

DIMS = synth.make_dims(33, 16, 24, 40)
B, T = 4, 21


def _model(device):
    from models.voicesplit.model import VoiceSplit
    m = VoiceSplit(config.AttrDict(synth.make_config_dict(DIMS)))
    m.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in synth.make_state_dict(DIMS, 5, "stress").items()})
    return m.to(device).train()


def _loss(mask, gw):
    return (mask * gw).sum() / mask.shape[0]          # a batch MEAN, like 20 - mean(snr) (generic_utils.py:473)


def _worker(rank, world, port, sync_bn, overlap, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from voicesplit_b200 import dist as vdist
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    d = vdist.init("nccl", dev)
    x, emb = synth.make_inputs(B, T, DIMS, 6)
    gw = np.random.default_rng(0).standard_normal((B, T, DIMS["num_freq"])).astype(np.float32)
    a, b = vdist.shard(B, rank, world)
    m = _model(dev).enable_data_parallel(d, sync_bn=sync_bn, overlap=overlap)
    mask = m(torch.from_numpy(x[a:b]).to(dev), torch.from_numpy(emb[a:b]).to(dev))
    _loss(mask, torch.from_numpy(gw[a:b]).to(dev)).backward()
    n = vdist.allreduce_gradients(m, d)
    torch.cuda.synchronize()
    res = {"n": n, "mask": mask.detach().cpu(), "grads": {k: p.grad.cpu() for k, p in m.named_parameters()},
           "running_mean": m.state_dict()["conv.2.running_mean"].cpu()}
    if rank == 0:
        # single-process references on this rank's GPU: the whole batch, and each shard on its own
        ref = _model(dev)
        full = ref(torch.from_numpy(x).to(dev), torch.from_numpy(emb).to(dev))
        _loss(full, torch.from_numpy(gw).to(dev)).backward()
        res["full_mask"] = full.detach().cpu()
        res["full_grads"] = {k: p.grad.cpu() for k, p in ref.named_parameters()}
        res["full_running_mean"] = ref.state_dict()["conv.2.running_mean"].cpu()
        shard_grads = []
        for r in range(world):
            sa, sb = vdist.shard(B, r, world)
            sm = _model(dev)
            _loss(sm(torch.from_numpy(x[sa:sb]).to(dev), torch.from_numpy(emb[sa:sb]).to(dev)), torch.from_numpy(gw[sa:sb]).to(dev)).backward()
            shard_grads.append({k: p.grad.cpu() for k, p in sm.named_parameters()})
        res["shard_mean_grads"] = {k: sum(g[k] for g in shard_grads) / world for k in shard_grads[0]}
    out.put((rank, res))
    d.barrier()
    d.destroy_process_group()


def _run(sync_bn, overlap):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, sync_bn, overlap, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=600) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    return got


def _close(a, b, rel=2e-3):
    scale = float(b.abs().max())
    return float((a - b).abs().max()) <= rel * scale + 1e-7


@pytest.mark.parametrize("overlap", [True, False])
def test_syncbn_dp_gradients_equal_the_single_process_gradient(overlap):
    got = _run(True, overlap)
    r0, r1 = got[0], got[1]
    assert r0["n"] == sum(v.numel() for v in r0["grads"].values())
    full_mask = r0["full_mask"]
    assert torch.allclose(torch.cat([r0["mask"], r1["mask"]]), full_mask, atol=2e-5)          # statistics of the concatenated batch
    assert torch.allclose(r0["running_mean"], r0["full_running_mean"], atol=1e-6)
    gmax = max(float(v.abs().max()) for v in r0["full_grads"].values())
    for k, g in r0["full_grads"].items():
        assert torch.equal(r0["grads"][k], r1["grads"][k]), k                                # both ranks hold the same reduced gradient
        assert float((r0["grads"][k] - g).abs().max()) <= 2e-3 * float(g.abs().max()) + 2e-6 * gmax, k


def test_per_rank_bn_dp_gradients_equal_the_mean_of_the_shard_gradients():
    got = _run(False, True)
    r0, r1 = got[0], got[1]
    gmax = max(float(v.abs().max()) for v in r0["shard_mean_grads"].values())
    for k, g in r0["shard_mean_grads"].items():
        assert torch.equal(r0["grads"][k], r1["grads"][k]), k
        assert float((r0["grads"][k] - g).abs().max()) <= 2e-4 * float(g.abs().max()) + 1e-6 * gmax, k
    # and they differ from the single-process gradient (different BatchNorm statistics): the documented BN caveat
    assert not _close(r0["grads"]["conv.1.weight"], r0["full_grads"]["conv.1.weight"], rel=1e-3)
