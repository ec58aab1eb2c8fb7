"""Generate tests/golden/*.npz from the UNMODIFIED reference (build container only).

    python tests/golden/make_golden.py

Imports /root/reference/proxmin under the alias ``proxmin_ref`` (tests/apis.py),
runs every case in tests/cases.py on it and stores the outputs.  The reference
cannot travel to the GPU box, these vectors do.  Library versions are recorded in
tests/golden/MANIFEST.json because BLAS summation order is part of the result.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import apis  # noqa: E402
import cases  # noqa: E402


def main():
    ref = apis.reference()
    if ref is None:
        raise SystemExit("reference not available at /root/reference")
    manifest = {"numpy": np.__version__, "cases": {}}
    try:
        import scipy
        manifest["scipy"] = scipy.__version__
        from threadpoolctl import threadpool_info
        manifest["blas"] = [{k: d.get(k) for k in ("internal_api", "version", "num_threads")} for d in threadpool_info()]
    except Exception:  # pragma: no cover
        pass
    for name, fn in cases.CASES.items():
        out = fn(ref)
        out = {k: np.asarray(v) for k, v in out.items()}
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        manifest["cases"][name] = {k: [str(v.dtype), list(v.shape)] for k, v in out.items()}
        print(name, {k: v.shape for k, v in out.items()})
    with open(os.path.join(HERE, "MANIFEST.json"), "w") as fh:
        json.dump(manifest, fh, indent=1)


if __name__ == "__main__":
    main()
