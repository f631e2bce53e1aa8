"""dgl.data.DGLDataset stand-in (see ../__init__.py).  TEST INFRASTRUCTURE ONLY."""
import os


class DGLDataset:
    """Mirrors the part of DGLDataset.__init__ the reference depends on
    (dxdata.py:172, 320-338): set _raw_dir/save_path, then load() when a cache
    exists, else process().  Real DGL also calls save() after process(); the
    reference tree is read-only so that step is skipped."""

    def __init__(self, name, url=None, raw_dir=None, save_dir=None, hash_key=(), force_reload=False, verbose=False):
        self._name = name
        self._raw_dir = raw_dir
        self._save_dir = save_dir if save_dir is not None else raw_dir
        self._force_reload = force_reload
        if not force_reload and self.has_cache():
            self.load()
        else:
            self.process()

    @property
    def name(self):
        return self._name

    @property
    def raw_dir(self):
        return self._raw_dir

    @property
    def save_dir(self):
        return self._save_dir

    @property
    def save_path(self):
        return os.path.join(self._save_dir, self._name)
