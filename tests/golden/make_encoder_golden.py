"""Golden vectors for the GE2E speaker encoder from the notebook's own UNMODIFIED classes.

    python tests/golden/make_encoder_golden.py          (build container only; needs /root/reference)

The classes LinearNorm and SpeakerEncoder are taken at run time from the reference's exported notebook
(notebooks/GE2E-Seungwonpark-ExtractSpeakerEmbedding-adaptado-para-openvoicefilter.py:52-85) by parsing the file and
executing only those two class definitions plus the hyper-parameter assignments (:34-41) - the rest of the notebook loads
CUDA checkpoints and walks a dataset.  Weights and mel inputs are regenerated from seeds (voicesplit_b200.synth); only the
outputs are stored."""
import ast
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from voicesplit_b200 import synth  # noqa: E402

NOTEBOOK = "/root/reference/notebooks/GE2E-Seungwonpark-ExtractSpeakerEmbedding-adaptado-para-openvoicefilter.py"
# name, flavour, weight seed, input seed, mel frames per utterance
CASES = [("encoder_default", "default", 31, 41, (301, 120, 80)), ("encoder_stress", "stress", 32, 42, (301, 95))]


def load_notebook_classes():
    tree = ast.parse(open(NOTEBOOK).read())
    keep = []
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name in ("LinearNorm", "SpeakerEncoder"):
            keep.append(node)
        elif isinstance(node, ast.Assign) and all(isinstance(t, ast.Name) for t in node.targets) and \
                node.targets[0].id in ("num_mels", "n_fft", "emb_dim", "lstm_hidden", "lstm_layers", "window", "stride"):
            keep.append(node)
    ns = {"torch": torch, "nn": torch.nn}
    exec(compile(ast.Module(body=keep, type_ignores=[]), NOTEBOOK, "exec"), ns)
    return ns


def main():
    ns = load_notebook_classes()
    for name, flavour, wseed, iseed, frames in CASES:
        enc = ns["SpeakerEncoder"](ns["num_mels"], ns["lstm_layers"], ns["lstm_hidden"], ns["window"], ns["stride"]).eval()
        sd = synth.make_encoder_state_dict(wseed, flavour)
        enc.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
        outs = []
        with torch.no_grad():
            for mel in synth.encoder_mel_inputs(iseed, frames):
                outs.append(enc(torch.from_numpy(mel)).numpy())
        out = os.path.join(ROOT, "tests", "golden", f"{name}.npz")
        np.savez_compressed(out, flavour=flavour, wseed=wseed, iseed=iseed, frames=np.array(frames), dvec=np.stack(outs).astype(np.float32),
                            hyper=np.array([ns[k] for k in ("num_mels", "lstm_layers", "lstm_hidden", "window", "stride", "emb_dim")]),
                            torch_version=torch.__version__)
        print(name, np.stack(outs).shape, "norms", [float(np.linalg.norm(o)) for o in outs])


if __name__ == "__main__":
    main()
