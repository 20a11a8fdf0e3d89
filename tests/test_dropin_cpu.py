"""CPU: with gtos_b200/dropin in front of the reference's script directory, the UNMODIFIED reference
generator.py builds its Generator out of the B200 modules and ends up with the same state_dict keys/shapes
as the all-reference model (so reference checkpoints load, generator/work.py:109).  Needs /root/reference
(build container only); skipped on the GPU box."""
import importlib
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/generator"
REFS = {"generator": "/root/reference/generator", "translator": "/root/reference/translator"}

SCRIPT = r'''
import sys, json, torch
mode = sys.argv[1]
if mode == "dropin":
    sys.path.insert(0, "%(root)s/gtos_b200/dropin")
    sys.path.insert(0, "%(root)s")
sys.path.append("%(ref)s")
import generator as G
class V:
    def __init__(s, n): s.size, s.padding_idx, s.unk_idx = n, 0, 1
    def idx2token(s, i): return "t%%d" %% i
vocabs = {k: V(n) for k, n in dict(concept=50, concept_char=30, relation=40, token=60, token_char=30, predictable_token=45).items()}
m = G.Generator(vocabs, 32, 300, 32, 300, [(3, 256)], 128, 128, 100, 64, 2, 128, 256, 8, 0.2, 1, 2, 2, None, "cpu")
print(json.dumps({"mods": [type(m.graph_encoder).__module__, type(m.relation_encoder).__module__, type(m.decoder).__module__,
                           type(m.snt_encoder).__module__, type(m.concept_encoder).__module__],
                  "sd": {k: list(v.shape) for k, v in m.state_dict().items()}}))
'''


def _run(mode, ref):
    out = subprocess.run([sys.executable, "-c", SCRIPT % dict(root=ROOT, ref=ref), mode], capture_output=True, text=True,
                         timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    return json.loads(out.stdout.strip().splitlines()[-1])


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("task", ["generator", "translator"])
def test_reference_generator_builds_from_dropin_modules(task):
    ours, ref = _run("dropin", REFS[task]), _run("reference", REFS[task])
    assert ours["mods"][:4] == ["gtos_b200.graph_transformer", "gtos_b200.encoder", "gtos_b200.decoder",
                                "gtos_b200.transformer"]
    assert ours["mods"][4] != "gtos_b200.encoder"          # TokenEncoder stays the reference's (out of scope)
    assert ours["sd"] == ref["sd"]
