"""Import the UNMODIFIED reference (/root/reference/{model,dxdata}.py) under the
dgl/mido stand-ins in oracle/shim.  TEST INFRASTRUCTURE ONLY.

Works only in the build container (where /root/reference is mounted).  It is
used to (a) validate oracle/dxvae_oracle.py against the real thing and (b)
generate the fixtures in tests/golden/ (oracle/make_golden.py).  Nothing that
runs on the GPU box may call this.
"""
import importlib.util
import os
import sys

REF_ROOT = os.environ.get("DXVAE_REFERENCE_ROOT", "/root/reference")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shim")


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "model.py"))


def _load(name):
    spec = importlib.util.spec_from_file_location("_dxvae_ref_" + name, os.path.join(REF_ROOT, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_reference():
    """Returns (model_module, dxdata_module) of the reference, run on CPU."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    os.environ.setdefault("CUDA_VISIBLE_DEVICES", "")  # model.py:13 auto-selects cuda
    if _SHIM not in sys.path:
        sys.path.insert(0, _SHIM)
    return _load("model"), _load("dxdata")


def load_dataset_graphs():
    """The 1024 graphs of DX_data/DXDataset.bin as shim graph objects, file order."""
    _, dxdata = load_reference()
    ds = dxdata.DXDataset(raw_dir=os.path.join(REF_ROOT, "DX_data"))
    graphs = ds.graphs[0] if isinstance(ds.graphs, tuple) else ds.graphs  # dxdata.py:335 stores the tuple
    return list(graphs)
