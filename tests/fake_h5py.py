"""An in-memory stand-in for the part of h5py the checkpoint tree format uses (h5py itself is not in this
image).  It keeps the h5py behaviours both writers and readers depend on:

* groups list their members in ALPHABETICAL order (h5py's default, not creation order);
* `special_dtype(vlen=str)` / `check_dtype(vlen=...)` work through NumPy dtype metadata, as in h5py;
* a variable-length string dataset reads back as `bytes` (scalar) or an object array of `bytes`
  whose dtype still carries the vlen tag (h5py >= 3);
* `create_dataset` rejects fixed-width unicode arrays ("No conversion path for dtype('<U..')");
* `create_dataset` / `create_group` make intermediate groups and refuse to replace an existing name.

Only tests and the golden generator import it: it pins the HDF5 LAYOUT (paths, `type` attributes, `arr{k}`
names, string handling) our writer produces against what the reference's `_savetree_hdf5` produces, and
lets each side's reader load the other side's tree."""
import numpy as np

_STORE = {}                      # filepath -> root Group


def reset():
    _STORE.clear()


def special_dtype(vlen=None):
    assert vlen is str
    return np.dtype("O", metadata={"vlen": str})


def check_dtype(vlen=None):
    md = getattr(vlen, "metadata", None)
    return md.get("vlen") if md else None


class Dataset:
    def __init__(self, data=None, dtype=None):
        if dtype is not None and check_dtype(vlen=dtype) is str:
            arr = np.asarray(data, dtype=object)
            self._vlen, self._value = True, np.array([str(s) for s in arr.ravel()], dtype=object).reshape(arr.shape)
        elif isinstance(data, str):
            self._vlen, self._value = True, np.array(data, dtype=object)
        else:
            arr = np.asarray(data) if dtype is None else np.asarray(data, dtype=dtype)
            if arr.dtype.kind == "U":
                raise TypeError(f"No conversion path for dtype: {arr.dtype!r}")
            if arr.dtype.kind == "O":
                raise TypeError("Object dtype dtype('O') has no native HDF5 equivalent")
            self._vlen, self._value = False, arr.copy()

    @property
    def shape(self):
        return self._value.shape

    @property
    def dtype(self):
        return special_dtype(vlen=str) if self._vlen else self._value.dtype

    def __getitem__(self, key):
        assert key == () or key is Ellipsis
        if self._vlen:
            if self._value.shape == ():
                return self._value.item().encode("utf-8")
            out = np.empty(self._value.shape, dtype=special_dtype(vlen=str))
            out[...] = np.vectorize(lambda s: s.encode("utf-8"), otypes=[object])(self._value)
            return out
        return self._value[()] if self._value.shape == () else self._value.copy()

    def describe(self):
        if self._vlen:
            return {"kind": "vlen_str", "shape": list(self._value.shape), "data": self._value.tolist()}
        return {"kind": "array", "dtype": self._value.dtype.str, "shape": list(self._value.shape),
                "data": self._value.tolist()}


class Group:
    def __init__(self):
        self._members = {}
        self.attrs = {}

    # -- path helpers ---------------------------------------------------------------------------
    def _walk(self, name, create=False):
        parts = [p for p in name.split("/") if p]
        node = self
        for p in parts[:-1]:
            if p not in node._members:
                if not create:
                    raise KeyError(name)
                node._members[p] = Group()
            node = node._members[p]
            if not isinstance(node, Group):
                raise KeyError(name)
        return node, parts[-1]

    def __contains__(self, name):
        try:
            parent, leaf = self._walk(name)
        except KeyError:
            return False
        return leaf in parent._members

    def __getitem__(self, name):
        parent, leaf = self._walk(name)
        if leaf not in parent._members:
            raise KeyError(f"Unable to open object (object '{leaf}' doesn't exist)")
        return parent._members[leaf]

    def __delitem__(self, name):
        parent, leaf = self._walk(name)
        del parent._members[leaf]

    def create_group(self, name):
        parent, leaf = self._walk(name, create=True)
        if leaf in parent._members:
            raise ValueError(f"Unable to create group (name already exists): {name}")
        parent._members[leaf] = Group()
        return parent._members[leaf]

    def require_group(self, name):
        if name in self:
            node = self[name]
            if not isinstance(node, Group):
                raise TypeError(f"Incompatible object (Dataset) already exists: {name}")
            return node
        return self.create_group(name)

    def create_dataset(self, name, data=None, dtype=None, **kwargs):
        parent, leaf = self._walk(name, create=True)
        if leaf in parent._members:
            raise ValueError(f"Unable to create dataset (name already exists): {name}")
        parent._members[leaf] = Dataset(data=data, dtype=dtype)
        return parent._members[leaf]

    # -- mapping protocol, alphabetical like h5py -----------------------------------------------
    def keys(self):
        return sorted(self._members)

    def values(self):
        return [self._members[k] for k in self.keys()]

    def items(self):
        return [(k, self._members[k]) for k in self.keys()]

    def __iter__(self):
        return iter(self.keys())

    def __len__(self):
        return len(self._members)

    def __repr__(self):
        return "<fake HDF5 group>"

    def describe(self, prefix=""):
        """Flat, JSON-ready description of everything below this group."""
        out = {}
        for k, node in self.items():
            path = f"{prefix}/{k}" if prefix else k
            if isinstance(node, Group):
                out[path] = {"kind": "group", "attrs": dict(node.attrs)}
                out.update(node.describe(path))
            else:
                out[path] = node.describe()
        return out


class File(Group):
    def __new__(cls, filepath, mode="r"):
        filepath = str(filepath)
        if mode == "r":
            if filepath not in _STORE:
                raise FileNotFoundError(filepath)
            return _STORE[filepath]
        assert mode == "a"
        if filepath not in _STORE:
            _STORE[filepath] = super().__new__(cls)
            Group.__init__(_STORE[filepath])
            _STORE[filepath]._path = filepath
            open(filepath, "ab").close()      # callers test os.path.exists
        return _STORE[filepath]

    def __init__(self, filepath, mode="r"):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def from_description(filepath, desc):
    """Rebuild a stored file from `Group.describe()` output (the committed fixture)."""
    root = File(filepath, "a")
    for path, d in desc.items():
        if d["kind"] == "group":
            g = root.require_group(path)
            g.attrs.update(d["attrs"])
    for path, d in desc.items():
        if d["kind"] == "vlen_str":
            if d["shape"] == []:
                root.create_dataset(path, data=d["data"])
            else:
                root.create_dataset(path, data=np.array(d["data"], dtype=object).reshape(d["shape"]),
                                    dtype=special_dtype(vlen=str))
        elif d["kind"] == "array":
            root.create_dataset(path, data=np.array(d["data"], dtype=np.dtype(d["dtype"])).reshape(d["shape"]))
    return root
