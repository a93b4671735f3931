"""Import shims that let the UNMODIFIED reference (/root/reference) run on CPU in the build container.

Test infrastructure only (used by tests/golden/make_golden.py and tests that are skipped when
/root/reference is absent).  Nothing here is imported by the product package.

Shims:
  1. torch.cuda.set_device -> no-op        (reference models/dino/dino.py:46 calls it at import)
     Tensor.cuda -> identity                (reference models/dino/dn_components.py:36)
  2. fake timm.models.layers                (imported by reference backbone.py via convnext/swin, never executed)
  3. fake MultiScaleDeformableAttention     (forward = the reference's own ms_deform_attn_core_pytorch,
                                             backward = autograd through it; the reference has no CPU kernel)
  4. stub addict.Dict / yapf                (util/slconfig.py:13-14)
  5. backbone.is_main_process -> False      (avoid torchvision weight download, backbone.py:118-120)
"""
import sys
import types
import os

import torch

REF_ROOT = os.environ.get("DTLR_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "models", "dino"))


def _install_fake_modules():
    # --- timm
    if "timm" not in sys.modules:
        timm = types.ModuleType("timm")
        timm_models = types.ModuleType("timm.models")
        timm_layers = types.ModuleType("timm.models.layers")

        def trunc_normal_(t, std=1.0, **kw):
            return torch.nn.init.trunc_normal_(t, std=std)

        class DropPath(torch.nn.Identity):
            def __init__(self, *a, **k):
                super().__init__()

        def to_2tuple(x):
            return (x, x) if not isinstance(x, (tuple, list)) else tuple(x)

        timm_layers.trunc_normal_ = trunc_normal_
        timm_layers.DropPath = DropPath
        timm_layers.to_2tuple = to_2tuple
        timm.models = timm_models
        timm_models.layers = timm_layers
        sys.modules["timm"] = timm
        sys.modules["timm.models"] = timm_models
        sys.modules["timm.models.layers"] = timm_layers

    # --- addict
    if "addict" not in sys.modules:
        addict = types.ModuleType("addict")

        class Dict(dict):
            def __init__(self, *args, **kwargs):
                super().__init__()
                for a in args:
                    if a is None:
                        continue
                    for k, v in (a.items() if isinstance(a, dict) else a):
                        self[k] = self._hook(v)
                for k, v in kwargs.items():
                    self[k] = self._hook(v)

            @classmethod
            def _hook(cls, v):
                if isinstance(v, dict) and not isinstance(v, cls):
                    return cls(v)
                if isinstance(v, (list, tuple)):
                    return type(v)(cls._hook(x) for x in v)
                return v

            def __getattr__(self, k):
                try:
                    return self[k]
                except KeyError:
                    raise AttributeError(k)

            def __setattr__(self, k, v):
                self[k] = v

            def __missing__(self, k):
                raise KeyError(k)

            def to_dict(self):
                out = {}
                for k, v in self.items():
                    out[k] = v.to_dict() if isinstance(v, Dict) else v
                return out

        addict.Dict = Dict
        sys.modules["addict"] = addict

    # --- yapf
    if "yapf" not in sys.modules:
        yapf = types.ModuleType("yapf")
        yapflib = types.ModuleType("yapf.yapflib")
        yapf_api = types.ModuleType("yapf.yapflib.yapf_api")
        yapf_api.FormatCode = lambda text, **kw: (text, False)
        yapf.yapflib = yapflib
        yapflib.yapf_api = yapf_api
        sys.modules["yapf"] = yapf
        sys.modules["yapf.yapflib"] = yapflib
        sys.modules["yapf.yapflib.yapf_api"] = yapf_api


def _install_msda_shim():
    """fake compiled module: forward/backward through the reference's pure-PyTorch core."""
    mod = types.ModuleType("MultiScaleDeformableAttention")

    def _core():
        from models.dino.ops.functions.ms_deform_attn_func import ms_deform_attn_core_pytorch
        return ms_deform_attn_core_pytorch

    def ms_deform_attn_forward(value, shapes, lsi, loc, w, im2col_step):
        return _core()(value, shapes.tolist(), loc, w)

    def ms_deform_attn_backward(value, shapes, lsi, loc, w, grad_output, im2col_step):
        with torch.enable_grad():
            v = value.detach().requires_grad_(True)
            l = loc.detach().requires_grad_(True)
            a = w.detach().requires_grad_(True)
            out = _core()(v, shapes.tolist(), l, a)
            gv, gl, ga = torch.autograd.grad(out, (v, l, a), grad_output)
        return gv, gl, ga

    mod.ms_deform_attn_forward = ms_deform_attn_forward
    mod.ms_deform_attn_backward = ms_deform_attn_backward
    sys.modules["MultiScaleDeformableAttention"] = mod


_loaded = False


def load_reference():
    """Returns the reference `models`, `util` top-level packages imported from REF_ROOT."""
    global _loaded
    if not available():
        raise RuntimeError("reference tree not found at %s" % REF_ROOT)
    if not _loaded:
        _install_fake_modules()
        _install_msda_shim()
        torch.cuda.set_device = lambda *a, **k: None
        torch.Tensor.cuda = lambda self, *a, **k: self
        if REF_ROOT not in sys.path:
            sys.path.insert(0, REF_ROOT)
        import models  # noqa: F401  (registers 'dino')
        import models.dino.backbone as bb
        bb.is_main_process = lambda: False
        _loaded = True
    import models
    import util
    return models, util


def ref_args(config="config/Latin_CTC.py", **overrides):
    """SLConfig -> argparse-like namespace the way finetuning.py:149-155 does it."""
    load_reference()
    from util.slconfig import SLConfig
    cfg = SLConfig.fromfile(os.path.join(REF_ROOT, config))
    d = cfg._cfg_dict.to_dict()
    ns = types.SimpleNamespace(**d)
    ns.device = "cpu"
    ns.dataset_file = "IAM"
    for k, v in overrides.items():
        setattr(ns, k, v)
    return ns


def build_reference_model(args):
    models, _ = load_reference()
    from models.registry import MODULE_BUILD_FUNCS
    build = MODULE_BUILD_FUNCS.get("dino")
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model, criterion, post = build(args)
    return model, criterion, post
