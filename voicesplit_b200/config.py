"""The reference's config surface: a JSON file with // comments read into an attribute dict
(/root/reference/utils/generic_utils.py:560-573).  Restated here so this repo's tools do not
depend on the reference tree; the reference's own load_config works with the modules as well."""
import json
import re


class AttrDict(dict):
    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.__dict__ = self


def load_config(path):
    with open(path, "r") as f:
        text = f.read()
    text = re.sub(r"\\\n", "", text)
    text = re.sub(r"//.*\n", "\n", text)
    cfg = AttrDict()
    cfg.update(json.loads(text))
    return cfg
