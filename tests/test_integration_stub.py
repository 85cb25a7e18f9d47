"""INTEGRATION.md section 3 shows the reference-side ctypes binding a maintainer would write.  The snippet is
extracted from the document itself: on CPU its call is checked against the header's prototype (argument count,
as round 1 shipped a 20-argument call to an 18-argument function), on the GPU it is executed against the oracle."""
import ast
import os
import re

import pytest
import torch

from _util import O, ROOT, case_inputs, t


def _snippet():
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", md, flags=re.S)
    hits = [b for b in blocks if "cvcl_text_encoder_fwd" in b and "ctypes.CDLL" in b]
    assert len(hits) == 1, "INTEGRATION.md must hold exactly one ctypes binding stub"
    return hits[0]


def test_stub_argument_count_matches_header():
    import multimodal_baby_b200 as cv
    tree = ast.parse(_snippet())
    calls = [n for n in ast.walk(tree) if isinstance(n, ast.Call) and isinstance(n.func, ast.Attribute)
             and n.func.attr == "cvcl_text_encoder_fwd"]
    assert len(calls) == 1
    want = len(cv._cabi.PROTOTYPES["cvcl_text_encoder_fwd"][1])
    assert len(calls[0].args) == want == 18
    # ... and the header declares the same number of parameters
    hdr = open(os.path.join(ROOT, "include", "cvcl_b200.h")).read()
    decl = re.search(r"int cvcl_text_encoder_fwd\((.*?)\);", hdr, flags=re.S).group(1)
    assert len([a for a in decl.split(",") if a.strip()]) == want


@pytest.mark.gpu
def test_stub_runs_and_matches_oracle():
    import multimodal_baby_b200 as cv
    cv._cabi.load()                                    # builds the library if needed
    ns = {}
    cwd = os.getcwd()
    os.chdir(ROOT)                                     # the snippet opens the library by its in-tree path
    try:
        exec(compile(_snippet(), "INTEGRATION.md#3", "exec"), ns)
        inp = case_inputs(11, 37, 512, "flat")
        ids, lens, table = t(inp["ids"], "cuda"), t(inp["lens"], "cuda"), t(inp["table"], "cuda")
        got = ns["text_features"](ids, lens, table)
        torch.cuda.synchronize()
    finally:
        os.chdir(cwd)
    ref = O.encode_text(t(inp["ids"]), t(inp["lens"]), t(inp["table"]), "flat", True)
    ref = ref[0] if isinstance(ref, tuple) else ref
    assert float((got.cpu() - ref).abs().max()) <= 2e-6
