"""Imports the UNMODIFIED reference modules from baseline/_ref (tools/install_ref.sh; falls back to /root/reference in the build
container).

TEST INFRASTRUCTURE (see oracle/__init__.py): used by tests/, by `bench.py --impl reference` and by bench.py's cpu_baseline leg only.

The reference is six flat script directories whose files import each other by bare name (`from multiTransformer import ...`,
`from models import ...`), so every load happens with that directory on sys.path and under a private module name.  `hot_path`
replaces the bare name `multiTransformer` for the duration of the import -- that is the one-line switch INTEGRATION.md describes:
the reference's own models.py / train.py then run on the B200 kernels.
"""
import contextlib
import importlib.util
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_CANDIDATES = [os.path.join(ROOT, 'baseline', '_ref'), '/root/reference/transformer']


def ref_root():
    for c in _CANDIDATES:
        if os.path.isfile(os.path.join(c, 'MFT', 'multiTransformer.py')):
            return c
    return None


def available():
    return ref_root() is not None


def _stub_matplotlib():
    if 'matplotlib' in sys.modules:
        return
    try:
        import matplotlib  # noqa: F401
        import matplotlib.pyplot  # noqa: F401
    except ImportError:
        import types
        m = types.ModuleType('matplotlib')
        p = types.ModuleType('matplotlib.pyplot')
        m.pyplot = p
        sys.modules['matplotlib'] = m
        sys.modules['matplotlib.pyplot'] = p


@contextlib.contextmanager
def _bare_names(directory, overrides):
    """`directory` first on sys.path, the bare names in `overrides` pre-seeded in sys.modules, everything restored afterwards."""
    saved = {k: sys.modules.get(k) for k in ('multiTransformer', 'models', 'datasets', 'train')}
    for k in saved:
        sys.modules.pop(k, None)
    for k, v in overrides.items():
        sys.modules[k] = v
    sys.path.insert(0, directory)
    try:
        yield
    finally:
        sys.path.remove(directory)
        for k, v in saved.items():
            sys.modules.pop(k, None)
            if v is not None:
                sys.modules[k] = v


_cache = {}


def load(model_dir, name, hot_path=None):
    """The reference module `name`.py of directory `model_dir` ('MFT', 'SFT', 'B2-Trans', 'B3-MFN').
    hot_path: a module to stand in for the bare name `multiTransformer` (None = the reference's own file)."""
    root = ref_root()
    if root is None:
        raise RuntimeError('reference sources not found: run tools/install_ref.sh in the build container (baseline/_ref is shipped by gpurun)')
    key = (model_dir, name, id(hot_path))
    if key in _cache:
        return _cache[key]
    _stub_matplotlib()
    d = os.path.join(root, model_dir)
    tag = 'dropin' if hot_path is not None else 'ref'
    overrides = {}
    if hot_path is not None:
        overrides['multiTransformer'] = hot_path
    with _bare_names(d, overrides):
        if name != 'multiTransformer' or hot_path is None:
            spec = importlib.util.spec_from_file_location(f'_{tag}_{model_dir.replace("-", "_")}_{name}', os.path.join(d, name + '.py'))
            mod = importlib.util.module_from_spec(spec)
            cwd = os.getcwd()
            with tempfile.TemporaryDirectory() as tmp:      # train.py opens ./train_cnn.log at import time
                os.chdir(tmp)
                try:
                    import warnings
                    with warnings.catch_warnings():
                        warnings.simplefilter('ignore')
                        spec.loader.exec_module(mod)
                finally:
                    os.chdir(cwd)
                    import logging
                    for h in list(logging.getLogger().handlers):      # ... and leaves a FileHandler into the deleted directory behind
                        if isinstance(h, logging.FileHandler):
                            logging.getLogger().removeHandler(h)
                            h.close()
        else:
            mod = hot_path
    _cache[key] = mod
    return mod


def cpu_instance(model):
    """Pin a reference model to the CPU inside a process that can see a GPU: the reference constructors default to cuda:0 when
    torch.cuda.is_available() (MFT/multiTransformer.py:177-179,284-286) and MultiCNNTransformer does not forward its `device` argument
    to the Transformer / MFN it builds (models.py:100) -- so move the instance and repoint every `.device` attribute."""
    import torch
    cpu = torch.device('cpu')
    model.to(cpu)
    for m in model.modules():
        if hasattr(m, 'device'):
            m.device = cpu
    return model
