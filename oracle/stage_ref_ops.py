"""ORACLE / BASELINE -- test infrastructure, NOT product code.

Stages the reference's own deformable-attention PYTHON caller code (models/dino/ops/functions + models/dino/ops/modules: the
`MSDeformAttnFunction` autograd function, `ms_deform_attn_core_pytorch` and the `MSDeformAttn` module -- two files, unmodified)
under oracle/_ref/ops/ so that the drop-in test can run them THROUGH dtlr_b200's operator boundary on the GPU box, where
/root/reference does not exist.  oracle/_ref/ is git-ignored (no reference source enters the repository's history) and travels
with the gpurun snapshot, exactly like the compiled reference CUDA op next to it.

    python oracle/stage_ref_ops.py          (build container only: needs /root/reference)
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("DTLR_REFERENCE", "/root/reference")
SRC = os.path.join(REF, "models", "dino", "ops")
OUT = os.path.join(HERE, "_ref", "ops")


def stage(force=False):
    if os.path.isdir(OUT) and not force:
        return OUT
    if not os.path.isdir(SRC):
        raise RuntimeError("reference sources not found under %s" % SRC)
    if os.path.isdir(OUT):
        shutil.rmtree(OUT)
    os.makedirs(OUT)
    for sub in ("functions", "modules"):
        shutil.copytree(os.path.join(SRC, sub), os.path.join(OUT, sub), ignore=shutil.ignore_patterns("__pycache__"))
    open(os.path.join(OUT, "__init__.py"), "w").close()
    return OUT


def load():
    """import the staged package as `ops` (after dtlr_b200.msda.install_as_reference_extension()); None when it was not staged"""
    if not os.path.isdir(OUT):
        return None
    root = os.path.dirname(OUT)
    if root not in sys.path:
        sys.path.insert(0, root)
    import importlib
    return importlib.import_module("ops.modules"), importlib.import_module("ops.functions.ms_deform_attn_func")


if __name__ == "__main__":
    print(stage(force="--force" in sys.argv))
