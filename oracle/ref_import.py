"""TEST / BENCH INFRASTRUCTURE (not product code): import the UNMODIFIED reference package for the drop-in tests
(tests/test_reference_driver.py) and for the reference arm of bench.py (`--impl reference`, `cpu_baseline`).

Search order: ``baseline/_ref`` (``pip install --no-deps --target baseline/_ref <copy of /root/reference>``: travels to the GPU box
with the snapshot) and then ``/root/reference/src`` (build container only).  The reference's I/O, plotting and hypertune modules
import h5py / tifffile / matplotlib / optuna / accelerate at module scope; none of them is touched by the hot path
(PtychoAD, CombinedLoss, CombinedConstraint, recon_step), so missing ones are replaced by inert stand-ins.
"""
import importlib
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))       # repo root (this file lives in oracle/)
CANDIDATES = [os.path.join(ROOT, "baseline", "_ref"), "/root/reference/src"]
_OPTIONAL = ["h5py", "tifffile", "optuna", "accelerate", "accelerate.utils", "matplotlib", "matplotlib.pyplot", "matplotlib.colors",
             "matplotlib.gridspec", "matplotlib.ticker", "matplotlib.patches", "mpl_toolkits", "mpl_toolkits.axes_grid1"]


class _Inert(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _Inert(k)

    def __call__(self, *a, **k):
        return _Inert("call")

    def __mro_entries__(self, bases):
        return (object,)


def reference_path():
    for p in CANDIDATES:
        if os.path.isdir(os.path.join(p, "ptyrad")):
            return p
    return None


def import_reference(with_driver: bool = False):
    """Returns a namespace with PtychoAD, CombinedLoss, CombinedConstraint (and recon_step if `with_driver`), or None."""
    p = reference_path()
    if p is None:
        return None
    if p not in sys.path:
        sys.path.insert(0, p)
    for n in _OPTIONAL:
        if n in sys.modules:
            continue
        try:
            importlib.import_module(n)
        except Exception:
            sys.modules[n] = _Inert(n)
    ns = types.SimpleNamespace(path=p)
    ns.PtychoAD = importlib.import_module("ptyrad.models").PtychoAD
    ns.CombinedLoss = importlib.import_module("ptyrad.losses").CombinedLoss
    ns.CombinedConstraint = importlib.import_module("ptyrad.constraints").CombinedConstraint
    if with_driver:
        ns.recon_step = importlib.import_module("ptyrad.reconstruction").recon_step
    return ns
