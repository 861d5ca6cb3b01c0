"""Import the UNMODIFIED reference hot-path modules from /root/reference.

TEST INFRASTRUCTURE ONLY.  Used in the build container to (a) validate the
restatement in ``oracle/mixstage_oracle.py`` and (b) generate the golden vectors
under ``tests/golden`` (see ``oracle/make_golden.py``).  Nothing on the product
path, in ``bench.py`` or in the ``-m gpu`` tests imports this file: the GPU box
has no ``/root/reference``.

Why a loader is needed: ``import model`` as a package pulls in
``model/trainer.py`` -> ``data`` -> h5py/librosa/nltk, none of which exist here.
We therefore register an *empty* package called ``model`` whose ``__path__`` is
the reference's ``src/model`` directory and load only the four files on the hot
path by name.  ``pycasper`` (github.com/chahuja/pycasper, un-pinned, un-vendored)
is replaced by a two-symbol stub whose semantics are *defined here* because the
real source is unavailable ("parity unpinned" for these two symbols):

* ``some_grad(module)``  -- context manager: ``requires_grad_(False)`` on entry for
  every parameter of ``module``, restored on exit.  Gradients still flow to the
  module *input* (reference call site joint_late_cluster_soft_style.py:198-200).
* ``LambdaScheduler(lambdas, **kw).step()`` -- returns the initial lambdas
  unchanged (constant lambda_D = lambda_gan = 1; call site gan.py:30-33,103).
"""
import contextlib
import importlib.util
import os
import sys
import types

REF_SRC = os.environ.get("MIXSTAGE_REFERENCE_SRC", "/root/reference/src")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_SRC, "model", "layers.py"))


def _install_pycasper_stub():
    if "pycasper.torchUtils" in sys.modules:
        return
    pkg = types.ModuleType("pycasper")
    tu = types.ModuleType("pycasper.torchUtils")

    @contextlib.contextmanager
    def some_grad(module):
        saved = [p.requires_grad for p in module.parameters()]
        for p in module.parameters():
            p.requires_grad_(False)
        try:
            yield
        finally:
            for p, r in zip(module.parameters(), saved):
                p.requires_grad_(r)

    class LambdaScheduler:
        def __init__(self, lambdas, **kwargs):
            self.lambdas = list(lambdas)

        def step(self):
            return list(self.lambdas)

    tu.some_grad = some_grad
    tu.LambdaScheduler = LambdaScheduler
    pkg.torchUtils = tu
    sys.modules["pycasper"] = pkg
    sys.modules["pycasper.torchUtils"] = tu


def load():
    """Returns a namespace with the reference classes (G, D, GAN, layers)."""
    if not available():
        raise RuntimeError("reference sources not found under %s" % REF_SRC)
    _install_pycasper_stub()
    name = "model"
    if name not in sys.modules or not getattr(sys.modules[name], "_mixstage_stub", False):
        pkg = types.ModuleType(name)
        pkg.__path__ = [os.path.join(REF_SRC, "model")]
        pkg._mixstage_stub = True
        sys.modules[name] = pkg
    mods = {}
    for sub in ("layers", "speech2gesture", "joint_late_cluster_soft_style", "gan"):
        full = "model." + sub
        if full not in sys.modules:
            spec = importlib.util.spec_from_file_location(
                full, os.path.join(REF_SRC, "model", sub + ".py"))
            m = importlib.util.module_from_spec(spec)
            sys.modules[full] = m
            spec.loader.exec_module(m)
        mods[sub] = sys.modules[full]
    ns = types.SimpleNamespace()
    ns.layers = mods["layers"]
    ns.G = mods["joint_late_cluster_soft_style"].JointLateClusterSoftStyle4_G
    ns.D = mods["joint_late_cluster_soft_style"].JointLateClusterSoftStyle4_D
    ns.GAN = mods["gan"].GAN
    ns.S2G = mods["speech2gesture"].Speech2Gesture_G
    return ns


def load_transform_functions():
    """The reference's own KMeans.get_feats / KMeans.predict / ZNorm.znorm / ZNorm.inv_znorm (src/data/transform.py), compiled
    from their source WITHOUT importing the module (its imports need h5py/librosa, absent here): the function definitions
    are cut out of the file with `ast` and executed unchanged; `self` is whatever object the caller passes."""
    import ast
    import torch
    path = os.path.join(REF_SRC, "data", "transform.py")
    src = open(path).read()
    tree = ast.parse(src)
    want = {("KMeans", "get_feats"), ("KMeans", "predict"), ("ZNorm", "znorm"), ("ZNorm", "inv_znorm")}
    out = {}
    for node in tree.body:
        if isinstance(node, ast.ClassDef):
            for fn in node.body:
                if isinstance(fn, ast.FunctionDef) and (node.name, fn.name) in want:
                    mod = ast.Module(body=[fn], type_ignores=[])
                    env = {"torch": torch}
                    exec(compile(mod, path, "exec"), env)
                    out[node.name + "." + fn.name] = env[fn.name]
    missing = [k for k in want if k[0] + "." + k[1] not in out]
    if missing:
        raise RuntimeError("reference functions not found: %s" % missing)
    return out


def load_metric_classes():
    """The reference's AverageMeter, L1, VelL1 and PCK classes (src/evaluation/metrics.py), compiled from their source
    without importing the module (its imports need the data pipeline)."""
    import ast
    import torch
    path = os.path.join(REF_SRC, "evaluation", "metrics.py")
    tree = ast.parse(open(path).read())
    env = {"torch": torch}
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name in ("AverageMeter", "L1", "VelL1", "PCK"):
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), env)
    return env
