"""One-off demonstration (SURVEY.md 8b "must run unchanged"): the reference's OWN training loop and validation
driver, unmodified, executed against this repo's module.

    reference code that runs:  train.train() (train.py:25-137) -> its model construction, Adam, validation() (utils/generic_utils.py:
                               476-529), the batch loop with `mask = model(mixed, emb); output = mixed * mask`, ap.torch_inv_spectrogram,
                               SiSNR_With_Pit, backward, optimizer.step, checkpoint save (+ validation again)
    what it gets from here:    models.voicesplit.model.VoiceSplit (import path resolves to THIS repo: repo root is first on sys.path),
                               the device audio processor as `ap`, synthetic loaders in the reference's collate formats

The reference tree cannot be installed (no setup.py) and does not exist on the GPU box: stage train.py, utils/ and config.json in
baseline/_ref/VoiceSplit/ (git-ignored, travels with gpurun) before the run:
    mkdir -p baseline/_ref/VoiceSplit && cp -r /root/reference/{train.py,utils,config.json} baseline/_ref/VoiceSplit/
Third-party packages the reference imports but this image lacks get inert stubs (librosa, soundfile, matplotlib, tensorboardX);
mir_eval.separation.bss_eval_sources is served by the engine's vs_sdr.  Log: profiles/r02_reference_drivers.txt."""
import os
import sys
import tempfile
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref", "VoiceSplit")
sys.path.insert(0, ROOT)
sys.path.append(REF)
import numpy as np
import torch

from voicesplit_b200 import synth
from voicesplit_b200.audio import DeviceAudioProcessor

assert os.path.isfile(os.path.join(REF, "train.py")), "stage the reference drivers first (see the docstring)"

_engine_box = {}


def _bss_eval_sources(ref, est, compute_permutation=False):
    eng = _engine_box["engine"]
    n = min(len(ref), len(est))
    sdr = eng.sdr(torch.as_tensor(np.asarray(ref[:n], np.float32))[None].cuda(), torch.as_tensor(np.asarray(est[:n], np.float32))[None].cuda())
    v = np.array([float(sdr[0])])
    return v, v, v, np.array([0])


for name in ("librosa", "librosa.util", "librosa.filters", "mir_eval", "mir_eval.separation", "soundfile", "matplotlib", "matplotlib.pylab",
             "tensorboardX"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["librosa"].util = sys.modules["librosa.util"]
sys.modules["librosa"].filters = sys.modules["librosa.filters"]
sys.modules["librosa"].__path__ = []                         # lets `from librosa.filters import mel` resolve the stub
sys.modules["librosa.filters"].mel = lambda *a, **k: None    # only the (unconfigured) WaveGlow back end would call these
sys.modules["librosa.util"].pad_center = lambda *a, **k: None
sys.modules["librosa.util"].tiny = lambda *a, **k: None
sys.modules["mir_eval.separation"].bss_eval_sources = _bss_eval_sources
sys.modules["matplotlib"].use = lambda *a, **k: None
sys.modules["matplotlib"].pylab = sys.modules["matplotlib.pylab"]


class _SummaryWriter:
    def __init__(self, *a, **k):
        self.scalars = []

    def add_scalar(self, tag, value, step):
        self.scalars.append((tag, float(value), int(step)))
        print(f"  [tensorboard] {tag} = {float(value):.4f} @ step {step}")

    def add_audio(self, *a, **k):
        pass

    def add_image(self, *a, **k):
        pass


sys.modules["tensorboardX"].SummaryWriter = _SummaryWriter

import train as ref_train                      # noqa: E402  the reference's train.py, unmodified
from utils.generic_utils import load_config    # noqa: E402  reference
import models.voicesplit.model as our_model    # noqa: E402

print("train.py from:", ref_train.__file__)
print("VoiceSplit from:", our_model.__file__)
assert our_model.__file__.startswith(ROOT) and "_ref" not in our_model.__file__
assert ref_train.VoiceSplit is our_model.VoiceSplit

c = load_config(os.path.join(REF, "config.json"))
c.train_config["epochs"] = 1
c.train_config["summary_interval"] = 1
c.train_config["checkpoint_interval"] = 2
c.train_config["learning_rate"] = 1e-3
audio = c.audio[c.audio["backend"]]
print("config: model", c.model, "num_freq", audio["num_freq"], "loss", c.loss["loss_name"])

dims = synth.make_dims(audio["num_freq"], c.model["emb_dim"], c.model["lstm_dim"], c.model["fc1_dim"], c.model["fc2_dim"])
T, L = 301, 48000


def item(seed):
    g = torch.Generator().manual_seed(seed)
    mixed_wav = synth.make_reference_audio(1, L, seed)[0] + 0.5 * synth.make_reference_audio(1, L, 100 + seed)[0]
    target_wav = synth.make_reference_audio(1, L, seed)[0]
    return mixed_wav.astype(np.float32), target_wav.astype(np.float32), torch.randn(c.model["emb_dim"], generator=g)


# a throw-away engine for the audio front end of the synthetic "dataset" (the model under training builds its own)
from voicesplit_b200.engine import MaskEngine  # noqa: E402
feng = MaskEngine(activation="mish", **dims)
feng.load_state_dict_tensors({k: torch.from_numpy(v).cuda() for k, v in synth.make_state_dict(dims, 0, "default").items() if v.dtype == np.float32})
_engine_box["engine"] = feng
ap = DeviceAudioProcessor(feng, audio)
ap.configure_training("q1")


def spec_and_angle(w):
    S, ph = ap.wav2spec(torch.from_numpy(w).cuda())
    return S.cpu(), torch.atan2(ph[..., 1], ph[..., 0]).cpu()


train_batches, test_items = [], []
for b in range(4):                             # 4 training batches of batch_size 2 (config.json:25), train_collate_fn format
    embs, tgts, mixs, lens, twavs, phs = [], [], [], [], [], []
    for i in range(2):
        mw, tw, emb = item(10 * b + i)
        ms, ang = spec_and_angle(mw)
        ts, _ = spec_and_angle(tw)
        embs.append(emb); tgts.append(ts); mixs.append(ms); lens.append(torch.tensor([L])); twavs.append(torch.from_numpy(tw)); phs.append(ang)
    train_batches.append(tuple(torch.stack(v) for v in (embs, tgts, mixs, lens, twavs, phs)))
for i in range(3):                             # eval_collate_fn format: a list holding one item tuple
    mw, tw, emb = item(500 + i)
    ms, ang = spec_and_angle(mw)
    ts, _ = spec_and_angle(tw)
    # validation() indexes `mixed_phase[0]` (generic_utils.py:502) although wav2spec returns the phase as [T, F]: with the shipped
    # backend that picks ONE frame, spec2wav then raises on the shape and the blanket `except: continue` (:522) drops every item.
    # Give the phase the leading axis that indexing expects, so that the reference's evaluation code actually executes.
    test_items.append([(emb, ts, ms, tw, mw, ang[None], torch.tensor([L]))])


class _ListLoader(list):
    pass


log_dir = tempfile.mkdtemp()
tb = ref_train.TensorboardWriter(log_dir, audio)
print("=== reference train.train() starts ===")
ref_train.train(None, log_dir, None, _ListLoader(train_batches), _ListLoader(test_items), tb, c, "voicesplit", ap, cuda=True)
print("=== reference train.train() returned ===")
losses = [v for tag, v, _ in tb.scalars if tag == "train_loss"]
print("train losses per step:", [round(v, 4) for v in losses])
assert len(losses) == 4 and all(np.isfinite(losses))
ck = [f for f in os.listdir(log_dir) if f.startswith("checkpoint_")]
print("checkpoints written by the reference loop:", ck)
assert ck, "checkpoint_interval = 2 must have produced a checkpoint"
ckpt = torch.load(os.path.join(log_dir, ck[0]), map_location="cpu")
m2 = our_model.VoiceSplit(c)
m2.load_state_dict(ckpt["model"])              # the reference's on-disk format loads back (test.py:43)
print("checkpoint keys:", sorted(ckpt.keys()), " step", ckpt["step"], " params", sum(p.numel() for p in m2.parameters()))
# the reference's test path: validation(test=True) over the loader (test.py:71)
from utils.generic_utils import validation, SiSNR_With_Pit  # noqa: E402
res = validation(SiSNR_With_Pit(), ap, m2.cuda(), _ListLoader(test_items), tb, 0, cuda=True, loss_name="si_snr", test=True)
print("reference validation(test=True) ->", res)
assert res is not None and np.isfinite(res[0]) and np.isfinite(res[1])
print("reference drivers OK")
