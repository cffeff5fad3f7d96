"""TEST INFRASTRUCTURE - import the UNMODIFIED reference modules from /root/reference.

Only usable in the build container (the GPU box has no /root/reference); used by
tests/golden/make_golden.py to generate the committed golden vectors and by the optional
tests that compare against the live reference.  librosa / mir_eval are absent from this image
and only needed by code outside the hot path, so inert stubs are registered for them
(SURVEY.md section 8c).
"""
import os
import sys
import types

REF_ROOT = "/root/reference"


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "models", "voicesplit", "model.py"))


def load():
    """Returns (VoiceSplit, VoiceFilter, generic_utils module) of the reference."""
    if not available():
        raise RuntimeError("reference tree not present")
    for name in ("librosa", "librosa.util", "mir_eval", "mir_eval.separation"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["mir_eval.separation"].bss_eval_sources = lambda *a, **k: None
    sys.modules["librosa"].util = sys.modules["librosa.util"]
    # the reference uses top-level package names 'models' and 'utils'; load it under private
    # names so it cannot shadow (or be shadowed by) this repo's own 'models' package
    import importlib.util

    def _load(modname, relpath):
        spec = importlib.util.spec_from_file_location(modname, os.path.join(REF_ROOT, relpath))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[modname] = mod
        spec.loader.exec_module(mod)
        return mod

    saved = {k: sys.modules.get(k) for k in ("utils", "utils.generic_utils")}
    try:
        gu = _load("_ref_generic_utils", "utils/generic_utils.py")
        pkg = types.ModuleType("utils"); pkg.generic_utils = gu
        sys.modules["utils"] = pkg; sys.modules["utils.generic_utils"] = gu
        vs = _load("_ref_voicesplit_model", "models/voicesplit/model.py")
        vf = _load("_ref_voicefilter_model", "models/voicefilter/model.py")
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return vs.VoiceSplit, vf.VoiceFilter, gu
