"""Import the *unmodified* reference (read-only at /root/reference) in this container.

Only used by ``make_golden.py`` (fixture generation) and by the optional
cross-check tests that skip themselves when /root/reference is absent (it does
not exist on the GPU box).  Nothing in the product imports this file.

The reference needs ``pytorch_lightning`` / ``shap`` / ``matplotlib`` / ``beir``
at import time; none is installed and none is used by the score-and-rank path,
so permissive stub modules are registered for them.  ``ranking()``
(scripts/ms_marco_eval.py:189-235) calls ``.cuda()`` and
``torch.cuda.synchronize()``; on a CPU-only box those two are patched to
no-ops *around the call* so the reference's own statements execute unmodified
on CPU tensors.
"""
import contextlib
import os
import sys
import types

# /root/reference in the build container; tests/golden/make_golden_cuda.py points this at the staged,
# git-ignored copy under baseline/_ref/ that travels to the GPU box
REFERENCE_ROOT = os.path.abspath(os.environ.get("CCR_REFERENCE_ROOT", "/root/reference"))


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "rime_lite"))


class _Anything:
    """Class usable as a base class, callable, attribute sink."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return self

    def __getattr__(self, name):
        return _Anything()


class _StubModule(types.ModuleType):
    __path__ = []  # behaves like a package

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything


_STUBS = [
    "pytorch_lightning",
    "pytorch_lightning.callbacks",
    "pytorch_lightning.callbacks.model_checkpoint",
    "pytorch_lightning.loggers",
    "pytorch_lightning.trainer",
    "pytorch_lightning.trainer.supporters",
    "shap",
    "shap.plots",
    "shap.plots._text",
    "matplotlib",
    "matplotlib.pyplot",
    "beir",
    "beir.retrieval",
    "beir.retrieval.evaluation",
    "beir.datasets",
    "beir.datasets.data_loader",
    # ccrec model/training modules pull in transformers+lightning training code that
    # the scoring path never touches; ms_marco_eval only imports names from them.
    "ccrec.models",
    "ccrec.models.bert_mt",
    "ccrec.models.bbpr",
    "ccrec.util.amazon_review_prime_pantry",
]


def _install_stubs():
    for name in _STUBS:
        if name not in sys.modules:
            sys.modules[name] = _StubModule(name)


def load_rime_lite_util():
    """-> the reference's ``rime_lite.util`` module (real code)."""
    _install_stubs()
    os.environ.setdefault("CCREC_INIT_ENV_DONE", "1")
    src = os.path.join(REFERENCE_ROOT, "src")
    if src not in sys.path:
        sys.path.insert(0, src)
    import rime_lite.util as ru  # noqa: E402

    return ru


def load_rime_lite_metrics():
    load_rime_lite_util()
    import rime_lite.metrics as rm

    return rm


def load_ms_marco_eval():
    """-> the reference's ``scripts/ms_marco_eval.py`` module (real code)."""
    load_rime_lite_util()
    scripts = os.path.join(REFERENCE_ROOT, "scripts")
    if scripts not in sys.path:
        sys.path.insert(0, scripts)
    import ms_marco_eval  # noqa: E402

    return ms_marco_eval


@contextlib.contextmanager
def cpu_as_cuda():
    """Make ``Tensor.cuda()`` / ``torch.cuda.synchronize()`` no-ops (CPU-only box)."""
    import torch

    if torch.cuda.is_available():
        yield
        return
    orig_cuda = torch.Tensor.cuda
    orig_sync = torch.cuda.synchronize
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.cuda.synchronize = lambda *a, **k: None
    try:
        yield
    finally:
        torch.Tensor.cuda = orig_cuda
        torch.cuda.synchronize = orig_sync
