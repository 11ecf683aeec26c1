"""Load the UNMODIFIED reference model from /root/reference (build container only).

TEST INFRASTRUCTURE (see oracle/__init__.py).  ``/root/reference`` does not exist on the GPU
box, so nothing that runs there may call this module; it is used by ``oracle/make_golden.py``
and by the CPU tests that are skipped when the reference tree is absent.

The reference's ``MedMamba.py`` imports two packages that are not installed here:

* ``timm.layers`` (``MedMamba.py:11``: ``DropPath``, ``trunc_normal_``) -- replaced by the two
  small stand-ins below (standard stochastic depth; torch's own truncated normal).
* ``mamba_ssm.ops.selective_scan_interface`` (``MedMamba.py:12``: ``selective_scan_fn``) --
  replaced by a function whose body is the reference's OWN TEXT of ``selective_scan_ref``:
  the docstring fragments at ``temp.py:57-139`` are pulled out of the file with ``ast`` at run
  time and compiled.  No line of it is restated here, so comparing it with
  ``oracle/selective_scan_ref.py`` pins our restatement to the reference's text.
"""
from __future__ import annotations

import ast
import importlib
import os
import sys
import textwrap
import types

REFERENCE_ROOT = os.environ.get("MEDMAMBA_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "MedMamba.py"))


def _scan_ref_from_reference_text():
    """Compile ``selective_scan_ref`` from the string fragments the reference keeps inside
    ``flops_selective_scan_ref`` (``temp.py:55-139``, each under an ``if False:`` guard)."""
    import torch
    import torch.nn.functional as F
    from einops import rearrange, repeat

    src = open(os.path.join(REFERENCE_ROOT, "temp.py")).read()
    tree = ast.parse(src)
    fn = next(n for n in tree.body
              if isinstance(n, ast.FunctionDef) and n.name == "flops_selective_scan_ref")
    fragments = []
    for node in fn.body:
        if isinstance(node, ast.If) and isinstance(node.test, ast.Constant) and node.test.value is False:
            for stmt in node.body:
                if (isinstance(stmt, ast.Expr) and isinstance(stmt.value, ast.Constant)
                        and isinstance(stmt.value.value, str)):
                    fragments.append(textwrap.dedent(stmt.value.value).strip("\n"))
    assert len(fragments) == 4, f"expected 4 text fragments in temp.py, found {len(fragments)}"
    body = "\n".join(fragments)
    code = ("def selective_scan_ref(u, delta, A, B, C, D=None, z=None, delta_bias=None,\n"
            "                       delta_softplus=False, return_last_state=False):\n"
            + textwrap.indent(body, "    ")
            + "\n    return out if not return_last_state else (out, last_state)\n")
    ns = dict(torch=torch, F=F, rearrange=rearrange, repeat=repeat)
    exec(compile(code, os.path.join(REFERENCE_ROOT, "temp.py") + ":57-139", "exec"), ns)
    return ns["selective_scan_ref"]


def _install_standins(scan_fn=None):
    import torch
    import torch.nn as nn

    if "timm.layers" not in sys.modules:
        class DropPath(nn.Module):
            """Stochastic depth per sample (stand-in for timm.layers.DropPath)."""

            def __init__(self, drop_prob: float = 0.0, scale_by_keep: bool = True):
                super().__init__()
                self.drop_prob = drop_prob
                self.scale_by_keep = scale_by_keep

            def forward(self, x):
                if self.drop_prob == 0.0 or not self.training:
                    return x
                keep = 1.0 - self.drop_prob
                mask = x.new_empty((x.shape[0],) + (1,) * (x.dim() - 1)).bernoulli_(keep)
                if keep > 0.0 and self.scale_by_keep:
                    mask.div_(keep)
                return x * mask

        def trunc_normal_(tensor, mean=0.0, std=1.0, a=-2.0, b=2.0):
            return nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)

        timm = types.ModuleType("timm")
        layers = types.ModuleType("timm.layers")
        layers.DropPath = DropPath
        layers.trunc_normal_ = trunc_normal_
        layers.to_2tuple = lambda v: (v, v) if not isinstance(v, tuple) else v
        timm.layers = layers
        sys.modules["timm"] = timm
        sys.modules["timm.layers"] = layers

    ref_scan = _scan_ref_from_reference_text()
    mamba = types.ModuleType("mamba_ssm")
    ops = types.ModuleType("mamba_ssm.ops")
    iface = types.ModuleType("mamba_ssm.ops.selective_scan_interface")
    iface.selective_scan_ref = ref_scan
    iface.selective_scan_fn = scan_fn if scan_fn is not None else ref_scan
    mamba.ops = ops
    ops.selective_scan_interface = iface
    sys.modules["mamba_ssm"] = mamba
    sys.modules["mamba_ssm.ops"] = ops
    sys.modules["mamba_ssm.ops.selective_scan_interface"] = iface
    return iface


def load_reference(scan_fn=None):
    """Return (MedMamba module, scan-interface stand-in module).  ``scan_fn`` overrides what the
    reference's ``selective_scan_fn`` name is bound to (default: the reference's own text)."""
    if not reference_available():
        raise FileNotFoundError(f"reference tree not found at {REFERENCE_ROOT}")
    iface = _install_standins(scan_fn)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    sys.modules.pop("MedMamba", None)
    mod = importlib.import_module("MedMamba")
    if scan_fn is not None:
        mod.selective_scan_fn = scan_fn
    return mod, iface


def reference_scan_ref():
    """The reference's own text of selective_scan_ref, compiled (container only)."""
    if not reference_available():
        raise FileNotFoundError(f"reference tree not found at {REFERENCE_ROOT}")
    return _scan_ref_from_reference_text()
