"""Uniform access to the three implementations the tests compare.

* ``reference()`` -- the unmodified reference, alias-imported from /root/reference
  (only exists in the build container; ``None`` elsewhere).
* ``oracle()``    -- ``oracle/proxmin_oracle.py`` wrapped to the reference's call
  signatures and return shapes.
* ``product()``   -- ``proxmin_b200`` (needs the CUDA library and a GPU to compute).

Each returned object exposes ``pgm adaprox admm sdmm bsdmm``, the ``prox_*``
operators, ``AlternatingProjections``, ``nmf`` (namespace with ``nmf``,
``grad_likelihood``, ``log_likelihood``, ``step_pgm``, ``step_adaprox``) and
``iterations()`` -> (iterations, sub_iterations) of the last solver call.
"""
import importlib.util
import logging
import os
import re
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REF_DIR = "/root/reference/proxmin"


class _LogTap(logging.Handler):
    """Parses 'Completed N iterations [and [a, b] sub-iterations]' from a logger."""

    rx = re.compile(r"Completed (\d+) iterations(?: and \[([^\]]*)\] sub-iterations)?")

    def __init__(self):
        super().__init__(level=logging.DEBUG)
        self.last = (None, None)

    def emit(self, record):
        m = self.rx.search(record.getMessage())
        if m:
            sub = [int(s) for s in m.group(2).split(",")] if m.group(2) else None
            self.last = (int(m.group(1)), sub)


def _tap(logger_name):
    lg = logging.getLogger(logger_name)
    for h in lg.handlers:
        if isinstance(h, _LogTap):
            return h
    tap = _LogTap()
    lg.addHandler(tap)
    if lg.level == logging.NOTSET or lg.level > logging.INFO:
        lg.setLevel(logging.INFO)
    return tap


def reference():
    if not os.path.isdir(REF_DIR):
        return None
    if "proxmin_ref" not in sys.modules:
        spec = importlib.util.spec_from_file_location(
            "proxmin_ref", os.path.join(REF_DIR, "__init__.py"), submodule_search_locations=[REF_DIR])
        mod = importlib.util.module_from_spec(spec)
        sys.modules["proxmin_ref"] = mod
        spec.loader.exec_module(mod)
    mod = sys.modules["proxmin_ref"]
    # the reference logs to "proxmin" -- same logger name as the product; tap it
    tap = _tap("proxmin")
    api = types.SimpleNamespace(name="reference", iterations=lambda: tap.last)
    for k in ("pgm", "adaprox", "admm", "sdmm", "bsdmm", "AlternatingProjections"):
        setattr(api, k, getattr(mod, k))
    for k in dir(mod.operators):
        if k.startswith("prox_"):
            setattr(api, k, getattr(mod.operators, k))
    api.nmf = mod.nmf
    api.utils = mod.utils
    return api


def oracle():
    from oracle import proxmin_oracle as o

    state = {"last": (None, None)}
    api = types.SimpleNamespace(name="oracle", iterations=lambda: state["last"])

    def pgm(*a, **k):
        c, G, S, n = o.pgm(*a, **k)
        state["last"] = (n, None)
        return c, G, S

    def adaprox(*a, **k):
        c, M, V, Vh, n, sub = o.adaprox(*a, **k)
        state["last"] = (n, list(sub))
        return c, M, V, Vh

    def admm(X, prox_f, step_f, prox_g=None, step_g=None, L=None, **k):
        c, e, n = o.admm(X, prox_f, step_f, prox_g=prox_g, step_g=step_g, L=L, **k)
        state["last"] = (n, None)
        return c, e

    def sdmm(X, prox_f, step_f, proxs_g=None, steps_g=None, Ls=None, **k):
        c, n = o.sdmm(X, prox_f, step_f, proxs_g=proxs_g, steps_g=steps_g, Ls=Ls, **k)
        state["last"] = (n, None)
        return c

    def bsdmm(X, proxs_f, steps_f_cb, proxs_g=None, **k):
        c, n = o.bsdmm(X, proxs_f, steps_f_cb, proxs_g=proxs_g, **k)
        state["last"] = (n, None)
        return c

    api.pgm, api.adaprox, api.admm, api.sdmm, api.bsdmm = pgm, adaprox, admm, sdmm, bsdmm
    api.AlternatingProjections = o.AlternatingProjections
    for k in dir(o):
        if k.startswith("prox_"):
            setattr(api, k, getattr(o, k))

    def nmf(Y, A, S, algorithm=None, **k):
        name = {None: "pgm", pgm: "pgm", adaprox: "adaprox", bsdmm: "bsdmm"}[algorithm]
        out = o.nmf(Y, A, S, algorithm=name, **k)
        if name == "pgm":
            state["last"] = (out[3], None)
            return out[:3]
        if name == "adaprox":
            state["last"] = (out[4], list(out[5]))
            return out[:4]
        state["last"] = (out[1], None)
        return out[0]

    api.nmf = types.SimpleNamespace(nmf=nmf, grad_likelihood=o.grad_likelihood,
                                    log_likelihood=o.log_likelihood, step_pgm=o.step_pgm,
                                    step_adaprox=o.step_adaprox)

    class Traceback(object):  # utils.py:104-116
        def __init__(self):
            self.trace = []

        def __call__(self, *X, it=None):
            self.trace.append(tuple(x.copy() for x in X))

    api.utils = types.SimpleNamespace(Traceback=Traceback, BarzilaiBorweinStepper=o.BarzilaiBorweinStepper)
    return api


def product():
    import proxmin_b200 as p

    tap = _tap("proxmin")
    api = types.SimpleNamespace(name="product", iterations=lambda: tap.last)
    for k in ("pgm", "adaprox", "admm", "sdmm", "bsdmm", "AlternatingProjections"):
        setattr(api, k, getattr(p, k))
    for k in dir(p.operators):
        if k.startswith("prox_"):
            setattr(api, k, getattr(p.operators, k))
    api.nmf = p.nmf
    api.utils = p.utils
    return api
